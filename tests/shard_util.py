"""Helpers for the sharded-scene tests: a per-rank simulator backed by the CPU oracle, and a 2-process gloo launcher."""
import os, socket, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_oracle_sim(scene, max_bodies):
    """Same interface nudge_b200.shard.ShardedSim expects from nudge_b200.Sim, on top of the CPU oracle (host memory is the state)."""
    from oracle import pyoracle
    from nudge_b200 import abi

    class OracleRank(pyoracle.OracleSim):
        def download_bodies(self): pass
        def upload_bodies(self): pass

        def apply_impulses(self, sweeps=1):
            for _ in range(sweeps):
                pyoracle.OracleSim.apply_impulses(self)

        def reload(self, new_scene):
            n = self.cache.count
            keep = (self.cache_tags[:n].copy(), self.cache_features[:n].copy(), self.cache_data[:n].copy())
            self._free()
            abi.HostState.__init__(self, new_scene, self.cap)
            self.cache_tags[:n], self.cache_features[:n], self.cache_data[:n] = keep
            self.cache.count = n

    return OracleRank(scene, contact_capacity=max(4096, 40 * max_bodies))


def make_gpu_sim(scene, max_bodies):
    import nudge_b200
    return nudge_b200.Sim(scene, max_bodies=max_bodies, max_boxes=max_bodies, max_spheres=(max_bodies if scene.n_spheres else 0), contact_capacity=max(4096, 40 * max_bodies))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, kind, steps, reshard_every, out_dir, n_boxes, init=None):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from nudge_b200 import scenes, shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = scenes.box_drop(n_boxes, iterations=4, seed=5, spacing=(2.4, 2.2, 2.4))
    if init is not None:            # start from a given state (transforms / momentum / idle of the global scene)
        st = np.load(init)
        g.transforms[:] = st["transforms"]; g.momentum[:] = st["momentum"]; g.idle[:] = st["idle"]
    if kind == "gpu_peer":          # the library's own push/pull kernels over CUDA-IPC peer memory (both ranks may share one GPU in the tests)
        sim = shard.ShardedSim(g, rank, world, make_gpu_sim, margin=0.5, transport="peer", nccl=False)
    else:
        sim = shard.ShardedSim(g, rank, world, make_oracle_sim if kind == "oracle" else make_gpu_sim, margin=0.5, transport="host")
    log = []
    for k in range(steps):
        if reshard_every and k and k % reshard_every == 0:
            sim.reshard()
        sim.step()
        sim.sim.download_bodies()
        log.append((sim.sim.transforms.copy(), sim.sim.momentum.copy(), sim.gids.copy(), sim.n_owned))
    final = sim.gather_global()
    np.savez(os.path.join(out_dir, "%s_rank%d.npz" % (kind, rank)), transforms=final.transforms, momentum=final.momentum,
             last_local_xf=log[-1][0], last_local_mom=log[-1][1], gids=log[-1][2], n_owned=log[-1][3],
             counts=np.array([sim.local_counts()["owned"], sim.local_counts()["ghosts"], sim.local_counts()["export"]]))
    dist.destroy_process_group()


def run_ranks(world, kind, steps=12, reshard_every=5, n_boxes=600, init=None):
    import torch.multiprocessing as mp
    out_dir = tempfile.mkdtemp(prefix="nb_shard_")
    port = free_port()
    mp.spawn(_worker, args=(world, port, kind, steps, reshard_every, out_dir, n_boxes, init), nprocs=world, join=True)
    return [np.load(os.path.join(out_dir, "%s_rank%d.npz" % (kind, r))) for r in range(world)]
