import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tests.test_device_sqp import PROBLEMS, callables, dense_pattern, pack_dense, _kernel
from tests.emu.emu import EmuSqp
name = sys.argv[1]
f, gf, ceq, jeq, cin, jin, x0, lb, ub, maxiter = PROBLEMS[name]
evalf, evalg, meq = callables(name)
n = len(x0); m = len(evalf(x0)[1])
colptr, prow = dense_pattern(n, m + 1)
k = _kernel(n, m, meq, colptr, prow, lb, ub, 1e-9, maxiter, 1)
emu = EmuSqp(n, m, meq, colptr, prow, lb, ub, 1e-9, maxiter, 1)
X = torch.from_numpy(np.clip(x0, lb, ub)[None].copy()).cuda()
Xe = np.clip(x0, lb, ub)[None].copy()
mode = np.zeros(1, dtype=np.int32)
k.start(1)
for rnd in range(4):
    xh = np.clip(X.cpu().numpy()[0], lb, ub)
    cs = np.concatenate([evalf(xh)[1], [evalf(xh)[0]]])[None]
    vs = pack_dense(evalg(xh)[1], evalg(xh)[0])[None]
    k.step(X, torch.from_numpy(cs).cuda(), torch.from_numpy(vs).cuda(), mode)
    xe = np.clip(Xe[0], lb, ub)
    emu.step(Xe, np.concatenate([evalf(xe)[1], [evalf(xe)[0]]])[None].copy(), pack_dense(evalg(xe)[1], evalg(xe)[0])[None].copy())
    sc = k.scalars(1)
    print("round", rnd, "gpu mode", mode, {a: float(b[0]) for a, b in sc.items()})
    print("         emu", {a: float(b) for a, b in emu.scalars(0).items()})
    print("   x gpu", X.cpu().numpy()[0], " x emu", Xe[0])
