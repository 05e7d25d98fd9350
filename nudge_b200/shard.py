"""A scene sharded across GPUs: slab decomposition along x, ghost bodies, one all-gather per solver sweep.

Every rank runs the complete, exact-order pipeline (collide ... advance) on its OWNED bodies plus GHOST copies of the
neighbours' bodies that lie within `halo` of its slab.  A contact between bodies of two ranks exists on both ranks; each
rank applies its impulses to both bodies, and after the warm start and after every sweep the ghost rows are overwritten
with their owners' BodyMomentum rows:  pack (nb_pack_momentum) -> ONE all-gather -> scatter (nb_unpack_momentum).
Within a rank the Gauss-Seidel order is exactly the reference's; across ranks the coupling is block-Jacobi, so results for
world_size > 1 differ from the single-GPU trajectory (they agree bit-for-bit for world_size == 1, and a rank's result is a
deterministic function of the partition, which is what the tests pin).  Ghosts are integrated locally with their owner's
momentum, so no transform exchange is needed inside a partition epoch; `reshard()` re-partitions from the gathered global
state (bodies that drift further than the halo margin must not happen between two reshards).

The partition logic is pure numpy and the exchange goes through torch.distributed (NCCL on device buffers, or gloo on host
buffers for the CPU tests), so the host logic is testable without a GPU."""
import numpy as np
from . import scenes as S


def slab_boundaries(x, world):
    """Equal-count slabs along x: world+1 boundaries, -inf / +inf at the ends."""
    qs = np.quantile(np.asarray(x, np.float64), np.linspace(0, 1, world + 1)[1:-1]) if world > 1 else np.zeros(0)
    return np.concatenate([[-np.inf], qs, [np.inf]])


def partition(x, world, halo):
    """x: positions of the dynamic bodies 1..N (index 0 = body 1).  Returns per rank (owned, ghosts, export) as arrays of global body ids.
    export[r] = owned bodies of r that some other rank holds as ghosts, in ascending id order: the rows r contributes to the all-gather."""
    x = np.asarray(x, np.float64)
    b = slab_boundaries(x, world)
    owner = np.clip(np.searchsorted(b, x, side="right") - 1, 0, world - 1)
    ids = np.arange(1, len(x) + 1, dtype=np.int64)
    owned, ghosts = [], []
    for r in range(world):
        mine = owner == r
        near = (~mine) & (x >= b[r] - halo) & (x < b[r + 1] + halo)
        owned.append(ids[mine]); ghosts.append(ids[near])
    exported = np.zeros(len(x) + 1, bool)
    for r in range(world):
        exported[ghosts[r]] = True
    export = [owned[r][exported[owned[r]]] for r in range(world)]
    return dict(boundaries=b, owner=owner, owned=owned, ghosts=ghosts, export=export)


def local_scene(g, owned, ghosts):
    """Local Scene of one rank: body 0 (static world) + owned + ghosts, and every collider (box or sphere) that sits on one of them,
    in the global collider order.  Colliders keep their GLOBAL tag, so every rank orders contacts the way the global scene would
    (box tags < sphere tags as in example/main.cpp:171)."""
    gids = np.concatenate([[0], owned, ghosts]).astype(np.int64)
    n = len(gids)
    lid = np.full(g.n_bodies, -1, np.int64); lid[gids] = np.arange(n)
    bb, sb = g.box_transforms["body"].astype(np.int64), g.sphere_transforms["body"].astype(np.int64)
    kb, ks = lid[bb] >= 0, lid[sb] >= 0
    s = S.Scene(n, int(kb.sum()), int(ks.sum()))
    s.name = g.name + "_shard"
    s.transforms[:] = g.transforms[gids]; s.properties[:] = g.properties[gids]; s.momentum[:] = g.momentum[gids]; s.idle[:] = g.idle[gids]
    s.box_data[:] = g.box_data[kb]; s.box_transforms[:] = g.box_transforms[kb]; s.box_tags[:] = g.box_tags[kb]
    s.box_transforms["body"] = lid[bb[kb]].astype(np.uint32)
    s.sphere_data[:] = g.sphere_data[ks]; s.sphere_transforms[:] = g.sphere_transforms[ks]; s.sphere_tags[:] = g.sphere_tags[ks]
    s.sphere_transforms["body"] = lid[sb[ks]].astype(np.uint32)
    for k in ("time_step", "iterations", "gravity", "damping"):
        setattr(s, k, getattr(g, k))
    return s, gids


def dataflow_plan(part, rank):
    """Plan of the experimental peer-memory exchange (include/nudge_b200.h, nb_exchange_*) for one rank, from the partition every
    rank computes identically.  Local body order is [world, owned..., ghosts...]; ghost j of a rank uses inbox slot j on that rank.
    Returns exp_off (CSR over local bodies), exp_rank, exp_slot (the subscribers of each owned body) and ghost_slot per local body."""
    world = len(part["owned"])
    owned, ghosts = part["owned"][rank], part["ghosts"][rank]
    n_local = 1 + len(owned) + len(ghosts)
    local_of = {int(g): 1 + k for k, g in enumerate(owned)}
    targets = [[] for _ in range(n_local)]
    for p in range(world):
        if p == rank:
            continue
        for j, g in enumerate(part["ghosts"][p]):
            k = local_of.get(int(g))
            if k is not None:               # I own this body: rank p wants it in its inbox slot j
                targets[k].append((p, j))
    exp_off = np.zeros(n_local + 1, np.uint32)
    exp_off[1:] = np.cumsum([len(t) for t in targets])
    flat = [t for ts in targets for t in ts]
    exp_rank = np.array([t[0] for t in flat], np.uint32); exp_slot = np.array([t[1] for t in flat], np.uint32)
    ghost_slot = np.full(n_local, 0xffffffff, np.uint32)
    ghost_slot[1 + len(owned):] = np.arange(len(ghosts), dtype=np.uint32)
    return dict(exp_off=exp_off, exp_rank=exp_rank, exp_slot=exp_slot, ghost_slot=ghost_slot)


class ShardedSim:
    """One rank of a sharded simulation.  `make_sim(scene, max_bodies)` builds the per-rank simulator (nudge_b200.Sim on the GPU box;
    the CPU oracle in the gloo tests); `comm` is a torch.distributed process group wrapper or None for world_size 1."""

    def __init__(self, global_scene, rank, world, make_sim, halo=8.0, device_exchange=False, dataflow=False):
        self.g = global_scene.copy()
        self.dataflow = bool(dataflow) and world > 1   # experimental: ghosts fed through peer-memory inboxes by the solver itself
        self.dataflow_ready = False
        self.rank, self.world, self.halo = rank, world, float(halo)
        self.make_sim = make_sim
        self.device_exchange = device_exchange
        self.sim = None
        self.exchange_rows = 0
        self._partition()

    # ---- partition bookkeeping (pure numpy) ----
    def _partition(self):
        g = self.g
        x = g.transforms["position"][1:, 0]
        self.part = partition(x, self.world, self.halo)
        owned, ghosts = self.part["owned"][self.rank], self.part["ghosts"][self.rank]
        scene, self.gids = local_scene(g, owned, ghosts)
        self.n_owned = len(owned)
        cap_bodies = int(1.5 * (len(x) / self.world + 2 * 4096) + 64)
        if self.sim is None:
            self.sim = self.make_sim(scene, max(cap_bodies, scene.n_bodies))
        else:
            self.sim.reload(scene)
        # exchange plan: where my export rows live locally, and for every ghost (owner rank, row in that rank's export list)
        exp = self.part["export"]
        self.max_export = max(1, max(len(e) for e in exp))
        lid = np.zeros(g.n_bodies, np.int64); lid[self.gids] = np.arange(len(self.gids))
        self.export_local = lid[exp[self.rank]].astype(np.uint32)
        pos_in_export = np.zeros(g.n_bodies, np.int64)
        for r in range(self.world):
            pos_in_export[exp[r]] = np.arange(len(exp[r]))
        owner_of = np.zeros(g.n_bodies, np.int64); owner_of[1:] = self.part["owner"]
        self.ghost_local = (1 + self.n_owned + np.arange(len(ghosts))).astype(np.uint32)
        self.ghost_source = (owner_of[ghosts] * self.max_export + pos_in_export[ghosts]).astype(np.uint32)
        self.exchange_rows = len(exp[self.rank])
        self._setup_buffers()
        if self.dataflow:
            self._setup_dataflow()

    def _setup_dataflow(self):
        import torch.distributed as dist
        sim = self.sim
        if not self.dataflow_ready:   # one inbox per rank for the lifetime of the simulator; handles swapped once
            cap = int(self.sim.cfg_max_bodies) if hasattr(self.sim, "cfg_max_bodies") else int(1.5 * (self.g.n_bodies / self.world + 2 * 4096) + 64)
            handle = sim.exchange_create(self.rank, self.world, cap, int(self.g.iterations) + 1)
            handles = [None] * self.world
            dist.all_gather_object(handles, handle)
            for p in range(self.world):
                sim.exchange_open(p, handles[p])
            self.dataflow_ready = True
        plan = dataflow_plan(self.part, self.rank)
        sim.exchange_plan(plan["exp_off"], plan["exp_rank"], plan["exp_slot"], plan["ghost_slot"])

    def _setup_buffers(self):
        import torch
        dev = "cuda" if self.device_exchange else "cpu"
        self.t_export = torch.zeros((self.max_export, 8), dtype=torch.float32, device=dev)
        self.t_gather = torch.zeros((self.world * self.max_export, 8), dtype=torch.float32, device=dev)
        if self.device_exchange:
            self.t_export_idx = torch.from_numpy(self.export_local.astype(np.int32)).to(dev)
            self.t_ghost_idx = torch.from_numpy(self.ghost_local.astype(np.int32)).to(dev)
            self.t_ghost_src = torch.from_numpy(self.ghost_source.astype(np.int32)).to(dev)

    # ---- ghost exchange: pack -> ONE all-gather -> unpack ----
    def exchange(self):
        if self.world == 1:
            return
        import torch.distributed as dist
        sim = self.sim
        if self.device_exchange:
            sim.pack_momentum(self.t_export_idx.data_ptr(), len(self.export_local), self.t_export.data_ptr())
            dist.all_gather_into_tensor(self.t_gather, self.t_export)
            sim.unpack_momentum(self.t_ghost_idx.data_ptr(), self.t_ghost_src.data_ptr(), len(self.ghost_local), self.t_gather.data_ptr())
        else:  # host path (gloo): same plan, numpy gathers
            sim.download_bodies()
            rows = sim.momentum.view(np.float32).reshape(-1, 8)
            self.t_export.numpy()[:len(self.export_local)] = rows[self.export_local.astype(np.int64)]
            parts = [self.t_gather[r * self.max_export:(r + 1) * self.max_export] for r in range(self.world)]
            dist.all_gather(parts, self.t_export)
            rows[self.ghost_local.astype(np.int64)] = self.t_gather.numpy()[self.ghost_source.astype(np.int64)]
            sim.upload_bodies()

    # ---- one simulation step (example/main.cpp:274-328) with the exchanges ----
    def step(self):
        if getattr(self, "graph", None) is not None:
            self.graph.replay()
            self.replayed_launches = getattr(self, "replayed_launches", 0) + self.graph_launches
            return
        self._step_launches()

    def launch_count(self):
        """Kernels the library launched for this rank, including the ones replayed from a recorded step."""
        return self.sim.launch_count() + getattr(self, "replayed_launches", 0)

    def capture(self, torch_stream):
        """Records one step (the library's launches on `torch_stream` - the stream the Sim was created with - and the NCCL
        all-gathers) into a CUDA graph; step() replays it until the next reshard().  Returns False (and keeps plain launches) if
        the capture is refused."""
        import torch
        self.graph = None
        if self.world > 1 and not self.device_exchange:
            return False
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            before = self.sim.launch_count()
            with torch.cuda.graph(g, stream=torch_stream):
                self._step_launches()
            torch.cuda.synchronize()
            self.graph_launches = self.sim.launch_count() - before
            self.graph = g
            return True
        except Exception as e:  # noqa: BLE001 - any capture failure means: stay on plain launches
            self.graph = None
            self.capture_error = repr(e)
            try: torch.cuda.synchronize()
            except Exception: pass
            return False

    def _step_launches(self):
        sim = self.sim
        if self.dataflow:
            sim.collide(); sim.apply_gravity_damping(); sim.read_cached_impulses(); sim.setup_contact_constraints_deferred()
            sim.solve_exchange(int(self.g.iterations))   # warm start + all sweeps, ghost hand-over inside the kernel
            sim.update_cached_impulses(); sim.write_cached_impulses(); sim.advance()
            return
        sim.collide(); sim.apply_gravity_damping(); sim.read_cached_impulses(); sim.setup_contact_constraints()
        self.exchange()
        for _ in range(int(self.g.iterations)):
            sim.apply_impulses()
            self.exchange()
        sim.update_cached_impulses(); sim.write_cached_impulses(); sim.advance()

    # ---- gather the global state on every rank and re-partition ----
    def gather_global(self):
        sim = self.sim
        sim.download_bodies()
        g = self.g
        own = slice(1, 1 + self.n_owned)
        ids = self.gids[own]
        if self.world == 1:
            g.transforms[ids] = sim.transforms[own]; g.momentum[ids] = sim.momentum[own]; g.idle[ids] = sim.idle[own]
            return g
        import torch, torch.distributed as dist
        payload = (ids.copy(), sim.transforms[own].copy(), sim.momentum[own].copy(), sim.idle[own].copy())
        out = [None] * self.world
        dist.all_gather_object(out, payload)
        for (i, t, m, c) in out:
            g.transforms[i] = t; g.momentum[i] = m; g.idle[i] = c
        return g

    def reshard(self):
        self.graph = None  # the partition (sizes, export plan) changes: a recorded step is stale
        self.gather_global()
        self._partition()

    def local_counts(self):
        return dict(owned=self.n_owned, ghosts=len(self.gids) - 1 - self.n_owned, export=self.exchange_rows)
