"""Generates tests/golden/steady_state.json by running the REFERENCE's own membrane ODE classes
(/root/reference/src/CGx/utils/membrane_ODE_systems.py, imported by path with stand-ins for the two modules it imports but
does not need for the computation: petsc4py -- only PETSc.Sys.Print -- and matplotlib.pyplot) on a stand-in problem object
that carries the constants of ProblemKNPEMI.setup_constants (KNPEMIx_problem.py:909-981) and given compartment sizes.
Run in the build container only (the GPU box has no /root/reference); the JSON is committed."""
import importlib.util
import json
import os
import sys
import time
import types

REF = "/root/reference/src/CGx/utils/membrane_ODE_systems.py"


def load_reference():
    petsc4py = types.ModuleType("petsc4py")
    petsc = types.ModuleType("petsc4py.PETSc")
    petsc.Sys = types.SimpleNamespace(Print=lambda *a, **k: None)
    petsc4py.PETSc = petsc
    sys.modules.setdefault("petsc4py", petsc4py)
    sys.modules.setdefault("petsc4py.PETSc", petsc)
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    spec = importlib.util.spec_from_file_location("ref_membrane_ode", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class V:
    def __init__(self, v):
        self.value = v


def stand_in_problem(geom, glia):
    p = types.SimpleNamespace()
    c = dict(R=8.314, F=96485.0, T=300.0, C_M=0.02, g_Na_bar=1200.0, g_K_bar=360.0, g_Na_leak=1.0, g_Na_leak_g=1.0,
             g_K_leak=4.0, g_K_leak_g=16.96, g_Cl_leak=0.25, g_Cl_leak_g=2.0, phi_rest=-0.065, phi_m_init=-0.070,
             Na_i_init=10.0, Na_e_init=145.0, K_i_init=130.0, K_e_init=3.0, Cl_i_init=5.0, Cl_e_init=134.0,
             phi_m_g_init=-0.085, Na_i_g_init=15.0, K_i_g_init=100.0, Cl_i_g_init=5.0)
    for k, v in c.items():
        setattr(p, k, V(v))
    for k, v in geom.items():
        setattr(p, k, v)
    return p, c


def main():
    mod = load_reference()
    # compartment sizes of a 20 um cube holding cells that fill 1/8 of it with a membrane area of 2.4e-9 m^2
    cases = {"neuron": (dict(vol_i_n=1.0e-15, vol_e=7.0e-15, area_g_n=2.4e-9), False),
             "neuron_glia": (dict(vol_i_n=1.0e-15, vol_i_g=0.6e-15, vol_e=6.4e-15, area_g_n=2.4e-9, area_g_g=1.8e-9), True)}
    out = {}
    for name, (geom, glia) in cases.items():
        p, consts = stand_in_problem(geom, glia)
        cls = mod.ThreeCompartmentMembraneODESystem if glia else mod.TwoCompartmentMembraneODESystem
        tic = time.perf_counter()
        sol = cls(p, stimulus_flag=False).solve_ode_system()
        out[name] = dict(constants=consts, geometry=geom, glia=glia, steady_state=[float(v) for v in sol],
                         seconds=time.perf_counter() - tic)
        print(name, out[name]["seconds"], out[name]["steady_state"])
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "steady_state.json")
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
