import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import example_trace
from tests.helpers import golden
os.environ["OGB200_MAX_OUTER"] = "2"
for tag in sys.argv[1:] or ["03", "11", "10"]:
    e = golden("example_" + tag)
    box, glb, text = example_trace.run_script(tag, intercept=False, exdir=example_trace.RUNDIR)
    prob = glb["prob"]
    eng = prob._engine
    lines = [l for l in text.splitlines() if "iteration" in l or "Exit mode" in l or "terminated" in l or "Iteration" in l or "incompatible" in l or "Singular" in l or "Positive" in l or "Current function" in l]
    print(tag, "launches", eng.launches, "\n   " + "\n   ".join(lines[:12]))
    x = np.clip(e["x0"], e["lb"], e["ub"])
    c, J = eng.eval_fd_host(x)
    print("   finite c/J at the shipped guess:", np.isfinite(c).all(), np.isfinite(J).all(), "p finite", np.isfinite(prob.p).all())
