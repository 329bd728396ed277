"""Time the on-device spectrum of a line probe against NumPy's FFT of the same samples on the host
(development aid for DESIGN.md; rows x frames as in config #3: 512 x 1000)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phonomena_b200.workloads import crystal_case

nx, ny, nz, steps = 512, 64, 32, 1000
e = crystal_case(nx, ny, nz).make_engine(dtype="f64", arith="fast", steps=steps)
pid = e.probe_add("uz", ny // 2, 0, steps)
l0 = e.launch_count
e.run(steps); e.sync()
win = np.hanning(steps)
for rep in range(2):
    t0 = time.perf_counter(); dev = np.abs(e.probe_dft_xt(pid, win, steps // 2, nx)) / np.sqrt(nx * steps); t_dev = time.perf_counter() - t0
t0 = time.perf_counter(); line = e.probe_read(pid); t_read = time.perf_counter() - t0
t0 = time.perf_counter(); ref = np.abs(np.fft.fft2(line * win, norm="ortho"))[:, :steps // 2]; t_np = time.perf_counter() - t0
print(json.dumps({"rows": nx, "frames": steps, "device_ms": t_dev * 1e3, "probe_read_ms": t_read * 1e3, "numpy_fft2_ms": t_np * 1e3,
                  "rel_l2": float(np.linalg.norm(dev - ref) / np.linalg.norm(ref)), "launches": e.launch_count - l0}))
