"""Side-by-side timing of the SpMV kernel variants on a bench workload (device-resident)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from feellgood_b200 import LinAlgebra, workloads

name = sys.argv[1] if len(sys.argv) > 1 else "film20m"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
w = workloads.build(name, scale=scale)
la = LinAlgebra(w.settings(), w.mesh)
la.set_state(w.u)
tm = w.timing()
la.step(w.Hext, tm, angle=0.3)
la.evolution()
la.step(w.Hext, tm, angle=0.4)
b = workloads.spmv_bytes(la.n, la.nnz)
ms = la.bench_spmv(40)
print("matrix-free: %.1f us  %.0f GB/s  (%.0f MB)" % (1e3 * ms, b / ms / 1e6, b / 1e6), flush=True)
la.set_operator("blocks")
la.step(w.Hext, tm, angle=0.5)
ms = la.bench_spmv(40)
print("assembled blocks: %.1f us  %.0f GB/s" % (1e3 * ms, workloads.spmv_bytes_blocks(la.n, la.nnz) / ms / 1e6))
