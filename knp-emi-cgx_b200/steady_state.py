"""Steady-state initial conditions: host-side mirror of the reference's membrane ODE systems.

When a config has no ``initial_conditions`` block (38 of the reference's 43 configs), ``ProblemKNPEMI.set_initial_conditions``
(KNPEMIx_problem.py:224-325) integrates a well-mixed compartment model -- neuron + ECS, or neuron + glia + ECS -- from the
default constants to rest and uses the result as initial membrane potential, concentrations and gates
(utils/membrane_ODE_systems.py: TwoCompartmentMembraneODESystem :585-827, ThreeCompartmentMembraneODESystem :118-475).
This is setup work on rank 0 of the host in the reference (scipy), and it is host work here too; the GPU path starts
after it.

State vector: [phi_m_n, Na_i_n, Na_e, K_i_n, K_e, Cl_i_n, Cl_e, (phi_m_g, Na_i_g, K_i_g, Cl_i_g,) n, m, h].
Every compartment exchanges ions with the ECS through its membrane: dc_i/dt = -I_k A / (z_k F V_i),
dc_e/dt = +I_k A / (z_k F V_e), C_M dphi_m/dt = -sum_k I_k, with the same channel models the facet kernel evaluates
(csrc/assembly.cu::facet_kernel) plus NKCC1, which is active here (utils/membrane_ODE_systems.py:104-115) although it is
the literal zero in the PDE forms.

Integration: scipy Radau with the reference's tolerances (rtol 1e-6, atol 1e-8), stopped when every derivative passes the
reference's test ``allclose(rhs, 0, rtol=1e-8, atol=1e-10)``.  The reference restarts the integrator every millisecond of
model time; here the restart interval doubles from 1 ms up to 10 s (same trajectory within the integrator tolerance, a few
hundred restarts instead of up to 5e5)."""
import numpy as np

Z = (1.0, 1.0, -1.0)                      # Na, K, Cl


def _f_nkcc1(K_e, K_e_0, K_min=3.0, eps=1e-6, cap=1.0):
    """Silencing factor of NKCC1 (utils/membrane_ODE_systems.py:104-115)."""
    if K_e <= K_min or K_e >= K_e_0:
        return 0.0
    val = 1.0 / (1.0 + (0.03 / max(K_e - K_e_0, eps)) ** 10)
    return min(max(val, 0.0), cap)


def _gate_rates(V):
    """Hodgkin-Huxley rate functions in 1/s, V = 1000 (phi_m - phi_rest) (KNPEMIx_ionic_model.py:443-470)."""
    an = 0.01e3 * (10.0 - V) / (np.exp((10.0 - V) / 10.0) - 1.0)
    bn = 0.125e3 * np.exp(-V / 80.0)
    am = 0.1e3 * (25.0 - V) / (np.exp((25.0 - V) / 10.0) - 1.0)
    bm = 4.0e3 * np.exp(-V / 18.0)
    ah = 0.07e3 * np.exp(-V / 20.0)
    bh = 1.0e3 / (np.exp((30.0 - V) / 10.0) + 1.0)
    return (an, bn), (am, bm), (ah, bh)


class MembraneSteadyState:
    """consts: dict with R, F, T, C_M, g_Na_bar, g_K_bar, g_leak (3), g_leak_g (3), phi_rest and the initial guesses
    phi_m, c_i (3), c_e (3), and for glia phi_m_g, c_i_g (3).  geom: vol_i_n, vol_e, area_n (+ vol_i_g, area_g)."""

    I_hat, P_Na_i, P_K_e = 0.25, 10.0, 1.5          # neuronal Na/K-ATPase
    S_KCC2, S_NKCC1 = 0.0068, 0.00023
    rho_pump = 1.1 * 1.12e-6                        # glial pump rate [mol / (m^2 s)]
    g_KCC1, g_NKCC1_g = 7e-2, 2e-2

    def __init__(self, consts, geom, glia=False):
        self.c, self.g, self.glia = consts, geom, glia
        self.psi = consts["R"] * consts["T"] / consts["F"]
        self.K_e_0 = consts["c_e"][1]
        # Kir reference potential: Nernst potential of the NEURONAL K+ guesses (utils/membrane_ODE_systems.py:272)
        self.E_K_0 = self.psi * np.log(consts["c_e"][1] / consts["c_i"][1])

    def nernst(self, k, ci, ce):
        return self.psi / Z[k] * np.log(ce / ci)

    def neuron_currents(self, phi, ci, ce, n, m, h):
        c = self.c
        E = [self.nernst(k, ci[k], ce[k]) for k in range(3)]
        I_atp = self.I_hat / ((1.0 + self.P_K_e / ce[1]) ** 2 * (1.0 + self.P_Na_i / ci[0]) ** 3)
        I_nkcc1 = self.S_NKCC1 * _f_nkcc1(ce[1], self.K_e_0) * np.log((ce[0] * ce[1] * ce[2] ** 2) / (ci[0] * ci[1] * ci[2] ** 2))
        I_kcc2 = self.S_KCC2 * np.log((ci[1] * ci[2]) / (ce[1] * ce[2]))
        return ((c["g_leak"][0] + c["g_Na_bar"] * m ** 3 * h) * (phi - E[0]) + 3.0 * I_atp - I_nkcc1,
                (c["g_leak"][1] + c["g_K_bar"] * n ** 4) * (phi - E[1]) - 2.0 * I_atp - I_nkcc1 + I_kcc2,
                c["g_leak"][2] * (phi - E[2]) + 2.0 * I_nkcc1 - I_kcc2)

    def glia_currents(self, phi, ci, ce):
        c, F = self.c, self.c["F"]
        E = [self.nernst(k, ci[k], ce[k]) for k in range(3)]
        I_pump = self.rho_pump * F / (1.0 + (self.P_Na_i / ci[0]) ** 1.5) / (1.0 + self.P_K_e / ce[1])
        I_nkcc1 = self.g_NKCC1_g * self.psi * _f_nkcc1(ce[1], self.K_e_0) * \
            np.log((ce[0] * ce[1] * ce[2] ** 2) / (ci[0] * ci[1] * ci[2] ** 2))
        I_kcc1 = self.g_KCC1 * self.psi * np.log((ci[1] * ci[2]) / (ce[1] * ce[2]))
        A = 1.0 + np.exp(0.433)
        B = 1.0 + np.exp(-(0.1186 + self.E_K_0) / 0.0441)
        Cc = 1.0 + np.exp(((phi - E[1]) + 0.0185) / 0.0425)
        D = 1.0 + np.exp(-(0.1186 + phi) / 0.0441)
        f_kir = A * B / (Cc * D) * np.sqrt(ce[1] / self.K_e_0)
        return (c["g_leak_g"][0] * (phi - E[0]) + 3.0 * I_pump - I_nkcc1,
                c["g_leak_g"][1] * f_kir * (phi - E[1]) - 2.0 * I_pump - I_nkcc1 + I_kcc1,
                c["g_leak_g"][2] * (phi - E[2]) + 2.0 * I_nkcc1 - I_kcc1)

    def rhs(self, t, x):
        c, g = self.c, self.g
        F, C_M = c["F"], c["C_M"]
        phi_n, ci_n, ce = x[0], (x[1], x[3], x[5]), (x[2], x[4], x[6])
        n, m, h = x[-3], x[-2], x[-1]
        In = self.neuron_currents(phi_n, ci_n, ce, n, m, h)
        out = np.zeros_like(x)
        out[0] = -(In[0] + In[1] + In[2]) / C_M
        for k in range(3):
            out[1 + 2 * k] = -In[k] / (Z[k] * F) * g["area_n"] / g["vol_i_n"]
            out[2 + 2 * k] = In[k] / (Z[k] * F) * g["area_n"] / g["vol_e"]
        if self.glia:
            phi_g, ci_g = x[7], (x[8], x[9], x[10])
            Ig = self.glia_currents(phi_g, ci_g, ce)
            out[7] = -(Ig[0] + Ig[1] + Ig[2]) / C_M
            for k in range(3):
                out[8 + k] = -Ig[k] / (Z[k] * F) * g["area_g"] / g["vol_i_g"]
                out[2 + 2 * k] += Ig[k] / (Z[k] * F) * g["area_g"] / g["vol_e"]
        for j, (a, b) in enumerate(_gate_rates((phi_n - c["phi_rest"]) * 1e3)):
            y = x[-3 + j]
            out[-3 + j] = a * (1.0 - y) - b * y
        return out

    def initial_state(self):
        c = self.c
        gates = [a / (a + b) for a, b in _gate_rates((c["phi_m"] - c["phi_rest"]) * 1e3)]
        x = [c["phi_m"], c["c_i"][0], c["c_e"][0], c["c_i"][1], c["c_e"][1], c["c_i"][2], c["c_e"][2]]
        if self.glia:
            x += [c["phi_m_g"], *c["c_i_g"]]
        return np.array(x + gates, dtype=np.float64)

    def solve(self, max_time=500.0, first_interval=1e-3, max_interval=10.0):
        """Integrates to rest.  Returns (state, model time, reached) -- `reached` False if max_time passed first."""
        from scipy.integrate import solve_ivp
        x, t, dt = self.initial_state(), 0.0, first_interval
        while t < max_time:
            sol = solve_ivp(self.rhs, [t, t + dt], x, method="Radau", rtol=1e-6, atol=1e-8)
            x, t = sol.y[:, -1], t + dt
            if not np.all(np.isfinite(x)):
                raise FloatingPointError("steady-state ODE: non-finite values in the solution")
            if np.allclose(self.rhs(t, x), 0.0, rtol=1e-8, atol=1e-10):
                return x, t, True
            dt = min(2.0 * dt, max_interval, max(max_time - t, first_interval))
        return x, t, False
