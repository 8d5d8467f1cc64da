#!/usr/bin/env python
"""PPO-style rollout over thousands of GPU environments (BASELINE.json configs[4]).

The batched counterpart of the reference's rollout loop
(/root/reference/baseline/PPO/train_PPO.py:78-107): observations never leave the GPU -- the
policy reads the environment's observation block through DLPack (zero copy), scores every
remaining net with the reference's own RepresentationNetwork (baseline/baseline_utils.py:231-379),
batched over nets and environments instead of a Python loop (xroute_env_b200/agent.py), samples one
legal net per environment, and steps the whole batch.  Only the N chosen
actions (4 bytes each) go to the host and N rewards come back.

    python examples/ppo_rollout.py --envs 1024 --steps 64            # one GPU's shard of 8192
    torchrun --nproc-per-node 8 examples/ppo_rollout.py --envs 1024  # 8192 environments
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xroute_env_b200 import VecGame, make_batch, preset_geometry  # noqa: E402
from xroute_env_b200.dist import allreduce_stats, shard_range  # noqa: E402


class NetScorer(nn.Module):
    """Policy / value heads on the reference's representation network, applied to all (environment, net)
    blocks at once (xroute_env_b200.agent.BatchedRepresentationNetwork: same parameters as
    baseline/baseline_utils.py:231-379, so a reference checkpoint can be loaded into ``self.rep``)."""

    def __init__(self):
        super().__init__()
        from xroute_env_b200.agent import BatchedRepresentationNetwork
        self.rep = BatchedRepresentationNetwork()
        self.pi = nn.Linear(128, 1)
        self.v = nn.Linear(64, 1)

    def forward(self, obs: torch.Tensor, n_remaining: torch.Tensor, max_nets: int):
        ob, rep, valid = self.rep(obs[:, :2 + 7 * max_nets], n_remaining)       # [N,64], [N,n,64], [N,n]
        logits = self.pi(torch.cat([rep, ob[:, None, :].expand_as(rep)], -1)).squeeze(-1)
        logits = logits.masked_fill(~valid, float("-inf"))
        return logits, self.v(ob).squeeze(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024, help="environments on this GPU")
    ap.add_argument("--nets", type=int, default=16)
    ap.add_argument("--preset", default="T1-1x1")
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--json", action="store_true", help="print one JSON line (used by bench.py)")
    ap.add_argument("--fp32", dest="tf32", action="store_false",
                    help="keep the policy's convolutions in full fp32 (default: cuDNN may use TF32 tensor cores, ~3x faster; "
                         "the policy samples its actions, the environment's arithmetic is integer either way)")
    ap.add_argument("--bf16", action="store_true", help="run the policy under bf16 autocast")
    ap.add_argument("--cudnn-benchmark", action="store_true", help="let cuDNN time its algorithms for the policy's shapes")
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    torch.backends.cudnn.allow_tf32 = args.tf32
    torch.backends.cuda.matmul.allow_tf32 = args.tf32
    torch.backends.cudnn.benchmark = args.cudnn_benchmark
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    geom = preset_geometry(args.preset)
    first, _ = shard_range(args.envs * world, rank, world)
    insts = make_batch(geom, args.envs, args.nets, seed=4242, first_env=first, max_degree=6)
    env = VecGame(geom, insts, device=local)
    policy = NetScorer().cuda().eval()
    env.reset()
    obs = env.obs_batch()                       # zero-copy, stays valid (updated in place) across steps
    buf_logp, buf_val, buf_rew, buf_done = [], [], [], []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ms_policy = ms_env = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env_steps = 0
    with torch.no_grad():
        for t in range(args.steps):
            n_rem = env.n_remaining.clone()
            live = n_rem > 0
            if not bool(live.any()):
                env.reset()
                continue
            ev[0].record()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
                logits, value = policy(obs, n_rem, args.nets)
            logits, value = logits.float(), value.float()
            logits[~live] = 0.0                  # finished environments idle (action 0)
            dist = torch.distributions.Categorical(logits=logits)
            pick = dist.sample()                 # rank among the remaining nets
            # rank -> net id: the order channel (channel 1) lists the remaining ids ascending
            order = obs[:, 1].flatten(1)[:, :args.nets]
            net_id = order.gather(1, pick[:, None]).squeeze(1).to(torch.int32)
            actions = torch.where(live, net_id, torch.zeros_like(net_id))
            ev[1].record()
            env.step(actions.cpu())              # N x 4 bytes to the host, the rest stays on the GPU
            ev[2].record()
            ev[2].synchronize()
            ms_policy += ev[0].elapsed_time(ev[1]); ms_env += ev[1].elapsed_time(ev[2])
            buf_logp.append(dist.log_prob(pick)); buf_val.append(value)
            buf_rew.append(env.reward.clone()); buf_done.append(env.done.clone())
            env_steps += int(live.sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = allreduce_stats(env.stats())
    rew = torch.stack(buf_rew)
    if rank == 0 and args.json:
        import json
        print(json.dumps({"grid": f"{args.preset} {geom.X}x{geom.Y}x{geom.Z}", "envs_per_gpu": args.envs, "n_gpus": world, "nets_per_env": args.nets,
                          "steps": args.steps, "value": env_steps * world / dt, "unit": "env-steps/s incl. policy",
                          "ms_policy_per_step": ms_policy / max(1, args.steps), "ms_env_per_step": ms_env / max(1, args.steps),
                          "policy_share": ms_policy / max(ms_policy + ms_env, 1e-9), "episodes_finished": stats["episodes"]}))
    elif rank == 0:
        print(f"{args.preset} {geom.X}x{geom.Y}x{geom.Z}: {args.envs} envs/GPU x {world} GPU, {args.nets} nets, "
              f"{args.steps} policy+env steps in {dt:.2f}s -> {env_steps * world / dt:.0f} env-steps/s incl. policy; "
              f"mean step reward {rew.mean().item():.1f}; episodes finished {stats['episodes']}")
    env.close()
    return env_steps


if __name__ == "__main__":
    main()
