"""Single-pair latency of the 1920-wide headline pair under the small-level variants, and that they change no bit:
PF_SOR_SMALL (single-region SOR with 16 warps, R = 2 / 4) and PF_FUSED_SMALL (64x4 assembly tiles on the coarse levels)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import pyflow
    from conftest import load_frame
    a, b = load_frame(1920, 1), load_frame(1920, 2)
    plan = pyflow.FlowPlan(1080, 1920, 3, mode="fp32_redblack", tuning=sys.argv[2])
    plan.upload(a, b)
    plan.solve(2)
    ms = plan.solve(5) / 5
    u, v, w = plan.download()
    np.savez(sys.argv[3], u=u, v=v, w=w)
    print("%.3f" % ms)
else:
    import numpy as np
    base = None
    for tune in ("latency", "throughput"):
        for small_sor, small_fused in ((0, 0), (1, 0), (0, 1), (1, 1)):
            env = dict(os.environ, PF_SOR_SMALL=str(small_sor), PF_FUSED_SMALL=str(small_fused))
            out = "/tmp/lv_%s_%d%d.npz" % (tune, small_sor, small_fused)
            r = subprocess.run([sys.executable, __file__, "child", tune, out], env=env, capture_output=True, text=True)
            if r.returncode:
                print(tune, small_sor, small_fused, "FAILED", r.stderr[-300:]); continue
            d = np.load(out)
            if base is None:
                base = d
            same = all(np.array_equal(d[k], base[k]) for k in ("u", "v", "w"))
            print("tuning %-10s PF_SOR_SMALL=%d PF_FUSED_SMALL=%d: %s ms/pair  %s" % (tune, small_sor, small_fused, r.stdout.strip(),
                  "bit-identical to the first run" if same else "DIFFERENT"), flush=True)
