#!/usr/bin/env python
"""PPO-style rollout over thousands of GPU environments (BASELINE.json configs[4]).

The batched counterpart of the reference's rollout loop
(/root/reference/baseline/PPO/train_PPO.py:78-107): observations never leave the GPU -- the
policy reads the environment's observation block through DLPack (zero copy), scores every
remaining net with a shared 3-D convolutional tower over its 7-channel block plus the obstacle
channel (the structure of the reference's RepresentationNetwork,
baseline/baseline_utils.py:231-379, batched over nets and environments instead of a Python
loop), samples one legal net per environment, and steps the whole batch.  Only the N chosen
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
    """Scores each remaining net from (obstacles, its 7 channels); value head on the mean."""

    def __init__(self, hidden: int = 16):
        super().__init__()
        self.tower = nn.Sequential(
            nn.Conv3d(8, hidden, 3, padding=1), nn.ReLU(),
            nn.Conv3d(hidden, hidden, 3, stride=2, padding=1), nn.ReLU(),
            nn.AdaptiveAvgPool3d(1), nn.Flatten())
        self.pi = nn.Linear(hidden, 1)
        self.v = nn.Linear(hidden, 1)

    def forward(self, obs: torch.Tensor, n_remaining: torch.Tensor, max_nets: int, chunk: int = 4096):
        # obs: [N, 2+7*max_nets, Z, Y, X] view of the library-owned buffer (read only)
        N, _, Z, Y, X = obs.shape
        nets = obs[:, 2:2 + 7 * max_nets].unflatten(1, (max_nets, 7))            # [N, n, 7, Z, Y, X] (view)
        valid = torch.arange(max_nets, device=obs.device)[None, :] < n_remaining[:, None]
        idx = valid.nonzero()                                                     # [M, 2] (env, rank)
        feats = []
        for s in range(0, idx.shape[0], chunk):
            e, r = idx[s:s + chunk, 0], idx[s:s + chunk, 1]
            x = torch.cat([obs[e, 0:1], nets[e, r]], 1)                          # gather only the live blocks
            feats.append(self.tower(x))
        feats = torch.cat(feats) if feats else obs.new_zeros((0, self.pi.in_features))
        logits = obs.new_full((N, max_nets), float("-inf"))
        logits[idx[:, 0], idx[:, 1]] = self.pi(feats).squeeze(-1)
        pooled = obs.new_zeros((N, feats.shape[1])).index_add_(0, idx[:, 0], feats)
        value = self.v(pooled / n_remaining.clamp(min=1)[:, None].float()).squeeze(-1)
        return logits, value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024, help="environments on this GPU")
    ap.add_argument("--nets", type=int, default=16)
    ap.add_argument("--preset", default="T1-1x1")
    ap.add_argument("--steps", type=int, default=64)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    geom = preset_geometry(args.preset)
    first, _ = shard_range(args.envs * world, rank, world)
    insts = make_batch(geom, args.envs, args.nets, seed=4242, first_env=first, max_degree=6)
    env = VecGame(geom, insts, device=local)
    policy = NetScorer().cuda().eval()
    env.reset()
    obs = env.obs_batch()                       # zero-copy, stays valid (updated in place) across steps
    rank_ids = torch.from_dlpack  # noqa: F841  (all views below come through DLPack as well)
    buf_logp, buf_val, buf_rew, buf_done = [], [], [], []
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
            logits, value = policy(obs, n_rem, args.nets)
            logits[~live] = 0.0                  # finished environments idle (action 0)
            dist = torch.distributions.Categorical(logits=logits)
            pick = dist.sample()                 # rank among the remaining nets
            # rank -> net id: the order channel (channel 1) lists the remaining ids ascending
            order = obs[:, 1].flatten(1)[:, :args.nets]
            net_id = order.gather(1, pick[:, None]).squeeze(1).to(torch.int32)
            actions = torch.where(live, net_id, torch.zeros_like(net_id))
            env.step(actions.cpu())              # N x 4 bytes to the host, the rest stays on the GPU
            buf_logp.append(dist.log_prob(pick)); buf_val.append(value)
            buf_rew.append(env.reward.clone()); buf_done.append(env.done.clone())
            env_steps += int(live.sum())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stats = allreduce_stats(env.stats())
    rew = torch.stack(buf_rew)
    if rank == 0:
        print(f"{args.preset} {geom.X}x{geom.Y}x{geom.Z}: {args.envs} envs/GPU x {world} GPU, {args.nets} nets, "
              f"{args.steps} policy+env steps in {dt:.2f}s -> {env_steps * world / dt:.0f} env-steps/s incl. policy; "
              f"mean step reward {rew.mean().item():.1f}; episodes finished {stats['episodes']}")
    env.close()
    return env_steps


if __name__ == "__main__":
    main()
