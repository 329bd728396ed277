"""CPU oracle (test infrastructure only).  See oracle/fdtd_numpy.py."""
