"""Study (not a test): how much of the sharded step's time is problem inflation (ghost bodies and their contacts) and how much is the
exchange?  One GPU settles the N x 65,536-box scene, cuts it into N cells with the production partition rule, and times plain nb_step
on ONE rank's local scene (owned + ghosts, no exchange) next to the 65,536-box scene itself.
  python tests/study_shard_inflation.py [N] > profiles/r02_shard_inflation.txt      (needs a GPU)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nudge_b200
from nudge_b200 import scenes, shard


def time_steps(sim, n=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        sim.step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for k in range(n):
        flush.fill_(k & 255)
        ev[k][0].record(); sim.step(); ev[k][1].record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    side = torch.cuda.Stream(); torch.cuda.set_stream(side)
    g = scenes.box_drop(65536 * world, iterations=8, seed=2)
    big = nudge_b200.Sim(g, stream=side.cuda_stream)
    for _ in range(900):
        big.step()
    big.download_bodies(); big.download_cache()
    g.transforms[:] = big.transforms; g.momentum[:] = big.momentum; g.idle[:] = big.idle
    print("global scene: %d bodies, %d contacts, %.3f ms per step on one GPU" % (g.n_bodies - 1, big.counts().contacts, time_steps(big)))
    part = shard.partition(g, world, margin=0.5)
    n = big.cache.count
    for r in sorted(set([0, world // 2, world - 1])):
        s, gids = shard.local_scene(g, part["owned"][r], part["ghosts"][r])
        loc = nudge_b200.Sim(s, stream=side.cuda_stream, contact_capacity=30 * s.n_bodies)
        loc.cache_tags[:n] = big.cache_tags[:n]; loc.cache_features[:n] = big.cache_features[:n]; loc.cache_data[:n] = big.cache_data[:n]; loc.cache.count = n
        if n <= loc.cap:
            loc.upload_cache()
        ms = time_steps(loc)
        c = loc.counts()
        print("rank %d of %d (grid %s): %d owned + %d ghosts, %d contacts, %d pairs: %.3f ms per plain nb_step (no exchange)" % (r, world, part["grid"], len(part["owned"][r]), len(part["ghosts"][r]), c.contacts, c.pairs, ms))
        loc.close()
    one = scenes.box_drop(65536, iterations=8, seed=2)
    o = nudge_b200.Sim(one, stream=side.cuda_stream)
    for _ in range(900):
        o.step()
    print("the 65,536-box scene itself: %d contacts, %.3f ms per step" % (o.counts().contacts, time_steps(o)))


if __name__ == "__main__":
    main()
