"""cProfile of the submitting thread in the end-to-end loop (what does one caldera_async() call cost on the host?)."""
import cProfile
import os
import pstats
import sys

sys.argv = [sys.argv[0], "--slots", "5", "--batch", "24", "--steps", "1", "--modes", "full"]
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import probe_e2e  # noqa: E402

probe_e2e.run("full", 2 * probe_e2e.nstreams)                      # every slot captures its graph
for rep in range(3):
    pr = cProfile.Profile()
    pr.enable()
    total, blocked, nblocked, steps = probe_e2e.run("full", 8 * probe_e2e.nstreams)
    pr.disable()
    print(f"profiled run {rep}: {8 * probe_e2e.nstreams / total:.1f} matrices/s, blocked {blocked:.3f} s of {total:.3f} s, "
          f"steps {[round(s, 3) for s in steps]}")
    if rep == 2 or max(steps[2:]) > 0.3:
        pstats.Stats(pr).sort_stats("tottime").print_stats(16)
