"""Phase timing of the frontier kernel (build with XR_NVCC_EXTRA=-DFR_TIMING): cycles of thread 0 per phase, summed over
all CTAs.    XR_NVCC_EXTRA=-DFR_TIMING python -m xroute_env_b200.build && python tools/diag_frontier.py [preset] [n_envs] [n_nets]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xroute_env_b200 import VecGame, make_batch, preset_geometry
preset = sys.argv[1] if len(sys.argv) > 1 else "SYN-256"
n_envs = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_nets = int(sys.argv[3]) if len(sys.argv) > 3 else 32
geom = preset_geometry(preset)
kw = dict(hot_spots=16, hot_sigma=32.0) if preset == "SYN-1024" else {}
insts = make_batch(geom, n_envs, n_nets, 20260000, **kw)
n_steps = int(sys.argv[4]) if len(sys.argv) > 4 else n_nets
rng = np.random.default_rng(1)
orders = np.stack([rng.permutation(i.net_ids) for i in insts], 1).astype(np.int32)
vg = VecGame(geom, insts, device=0, obs_max_nets=8 if preset == 'SYN-1024' else -1)
vg.reset()
rec = (C.c_uint64 * (8 * n_envs))()
worst = []
allrec = []
for t in range(n_steps):
    vg.step(orders[t])
    vg._L.xr_debug_env_records(vg._h, rec)
    r = np.array(rec[:], np.int64).reshape(n_envs, 8)
    acts = orders[t]
    for e in range(n_envs):
        pe = len(set(insts[e].ap_pin[insts[e].ap_net == acts[e]].tolist()))
        allrec.append([pe] + r[e].tolist())
    k = int(np.argmax(r[:, 0]))
    pins = len(set(insts[k].ap_pin[insts[k].ap_net == acts[k]].tolist()))
    worst.append((int(r[k, 0]), t, k, pins, r[k].tolist()))
torch.cuda.synchronize()
print("slowest net of every step: cycles | step env pins | rounds expanded conns expand_cyc classify_cyc refill_cyc refills max_open")
for w in sorted(worst, reverse=True)[:12]:
    print(f"   {w[0]:9d} | {w[1]:2d} {w[2]:2d} {w[3]:2d} | {w[4][1]:4d} {w[4][2]:7d} {w[4][3]:3d} {w[4][4]:9d} {w[4][5]:9d} {w[4][6] & ((1 << 40) - 1):9d} {w[4][6] >> 40:4d} {w[4][7]:6d}")
A = np.array(allrec, np.float64)
print("by pin count: nets | mean cycles | max cycles | rounds | expanded | conns | expand cyc | classify cyc | cycles/round | cycles/expanded")
for lo, hi in ((2, 3), (4, 7), (8, 12), (13, 99)):
    m = (A[:, 0] >= lo) & (A[:, 0] <= hi)
    if m.sum():
        q = A[m]
        print(f"   pins {lo:2d}-{hi:2d}: {int(m.sum()):4d} | {q[:, 1].mean():9.0f} | {q[:, 1].max():9.0f} | {q[:, 2].mean():6.1f} | {q[:, 3].mean():8.0f} | {q[:, 4].mean():5.1f} | "
              f"{q[:, 5].mean():9.0f} | {q[:, 6].mean():9.0f} | {q[:, 1].sum() / q[:, 2].sum():7.0f} | {q[:, 1].sum() / q[:, 3].sum():6.1f}")
print(f"mean of the per-step maxima: {np.mean([w[0] for w in worst]):.0f} cycles")
out = (C.c_uint64 * 16)()
vg._L.xr_debug_counters(vg._h, out)
v = [int(x) for x in out]
names = ["seed+connect", "boxes+push", "classify", "expand", "target", "walk", "commit"]
nets = max(v[13], 1)
print(f"{preset} x {n_envs} x {n_nets}: nets {v[13]}, connections {v[10]}, rounds {v[9]} (max per net {v[12]}), expanded entries {v[11]}")
print(f"kernel cycles per net: mean {v[0] / nets:.0f}, max {v[1]}; CTA 0 (most pins) mean {v[14] / n_steps:.0f} cycles, {v[15] / n_steps:.1f} rounds")
for k, n in enumerate(names):
    print(f"   {n:14s} {v[2 + k] / nets:10.0f} cycles/net  {100.0 * v[2 + k] / max(v[0], 1):5.1f} %   per connection {v[2 + k] / max(v[10], 1):8.0f}   per round {v[2 + k] / max(v[9], 1):8.0f}")
vg.close()
