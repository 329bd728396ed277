"""phonomena_b200 -- B200 (sm_100a) FDTD time-stepping engine for Phonomena.

Drop-in for the solver plugin interface of phonomena/simulation/solvers/*.py
(`Solver.init/run/cancel/test`); see solver_b200.py, include/phb200.h, INTEGRATION.md.
"""
__version__ = "0.1.0"
