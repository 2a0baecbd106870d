"""Phase cycle counters of the warp-specialised rollout (library built with -DNFSP_TQ_PROF).

    make -C <pkg>/csrc clean all EXTRA=-DNFSP_TQ_PROF && python profiles/tq_profile.py [patience_avg,patience_br]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import nfsp_b200  # noqa: E402
from nfsp_b200._lib import lib  # noqa: E402

n, T = 1 << 20, 8
sp = nfsp_b200.SelfPlay(n, seed=1234, rl_capacity=1 << 25, sl_capacity=1 << 23, max_steps_per_call=T, variant="tcgen05_ws")
if len(sys.argv) > 1:
    a, b = (int(v) for v in sys.argv[1].split(","))
    lib().nfsp_rollout_tune(sp.env._h, a, b)
reps = 5
for _ in range(reps):
    sp.rollout(T, insert=False)
    sp.counts.zero_()
torch.cuda.synchronize()
out = (C.c_uint64 * 24)()
lib().nfsp_rollout_profile(sp.env._h, out)
v = [int(x) for x in out]
sms, envw, epiw = 148, 23, 8
L = reps
e, sc, ep = v[0:8], v[8:16], v[16:24]
print("per launch and SM (cycles):")
print("  env warp: total %.0f  begin (ticket+row) %.0f  waiting %.0f  | set-steps/warp %.1f -> per set-step: begin %.0f wait %.0f other %.0f" % (
    e[0] / L / sms / envw, e[1] / L / sms / envw, e[2] / L / sms / envw, e[3] / L / sms / envw,
    e[1] / max(e[3], 1), e[2] / max(e[3], 1), (e[0] - e[1] - e[2]) / max(e[3], 1)))
print("  scheduler: total %.0f  tiles %.1f  rows/tile %.1f  loop iterations %.0f | per tile: TMEM-slot wait %.0f  fences+meta %.0f  MMA issue+commit %.0f  rest %.0f" % (
    sc[0] / L / sms, sc[2] / L / sms, sc[3] / max(sc[2], 1), sc[6] / L / sms, sc[1] / max(sc[2], 1), sc[5] / max(sc[2], 1), sc[4] / max(sc[2], 1),
    (sc[0] - sc[1] - sc[4] - sc[5]) / max(sc[2], 1)))
print("  epilogue warp: total %.0f  waiting %.0f  tiles %.1f  computed %.1f -> busy cycles per computed tile-warp %.0f" % (
    ep[0] / L / sms / epiw, ep[1] / L / sms / epiw, ep[2] / L / sms / epiw, ep[3] / L / sms / epiw, (ep[0] - ep[1]) / max(ep[3], 1)))
