"""Run the randomised parity tests over many more seeds than the test suite does: python tools/fuzz_soak.py [n]"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import pytest  # noqa: F401
import test_k1_fuzz_gpu as K1
import test_sed_fuzz_gpu as K4
import test_k2_fuzz_gpu as K2
import test_ensemble_fuzz_gpu as K3

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bad = 0
for seed in range(16, 16 + n):
    for name, fn in (("k1", K1.test_random_modelsets_match_the_oracle),
                     ("k4", K4.test_k4_random_modelsets_match_the_oracle), ("k2", K2.test_k2_random_filter_sets),
                     ("k3", K3.test_ensemble_on_random_modelsets)):
        try:
            fn(seed)
        except BaseException as e:          # pytest.skip raises a BaseException subclass
            if type(e).__name__ == "Skipped":
                continue
            bad += 1
            print("FAIL", name, "seed", seed, type(e).__name__, str(e)[:300])
            traceback.print_exc(limit=2)
print("soak over %d seeds: %d failures" % (n, bad))
