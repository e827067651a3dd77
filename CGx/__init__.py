"""Drop-in shim: the reference's import paths (``from CGx.KNPEMI.KNPEMIx_solver import SolverKNPEMI`` ...)
resolved to the B200-native implementation in ``knp-emi-cgx_b200/``."""
