"""Development aid: the fuzz test of tests/test_gpu_parity.py over many more seeds."""
import sys, traceback
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gpu_parity as t
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
for seed in range(lo, hi):
    for kind in ("sift", "superpoint"):
        try:
            t.test_fuzz_batched_paths_vs_reference_kernels.__wrapped__(kind, seed) if hasattr(t.test_fuzz_batched_paths_vs_reference_kernels, "__wrapped__") else t.test_fuzz_batched_paths_vs_reference_kernels(kind, seed)
        except Exception:
            bad += 1
            print("FAIL", kind, seed); traceback.print_exc(limit=2)
print("fuzz done: seeds %d..%d, failures %d" % (lo, hi, bad))
