"""TEST INFRASTRUCTURE ONLY: ctypes bindings for oracle/_ref (the unmodified reference compiled in place).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Structures mirror nudge.h:29-129 byte for byte (the reference's own uint16 index layout)."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Arena(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t)]


class ContactData(C.Structure):
    _fields_ = [("data", C.c_void_p), ("bodies", C.c_void_p), ("tags", C.c_void_p), ("capacity", C.c_uint32), ("count", C.c_uint32),
                ("sleeping_pairs", C.c_void_p), ("sleeping_count", C.c_uint32)]


class _Shapes(C.Structure):
    _fields_ = [("tags", C.c_void_p), ("data", C.c_void_p), ("transforms", C.c_void_p), ("count", C.c_uint32)]


class ColliderData(C.Structure):
    _fields_ = [("boxes", _Shapes), ("spheres", _Shapes)]


class BodyData(C.Structure):
    _fields_ = [("transforms", C.c_void_p), ("properties", C.c_void_p), ("momentum", C.c_void_p), ("idle_counters", C.c_void_p), ("count", C.c_uint32)]


class BodyConnections(C.Structure):
    _fields_ = [("data", C.c_void_p), ("count", C.c_uint32)]


class ContactCache(C.Structure):
    _fields_ = [("tags", C.c_void_p), ("data", C.c_void_p), ("capacity", C.c_uint32), ("count", C.c_uint32)]


class ActiveBodies(C.Structure):
    _fields_ = [("indices", C.c_void_p), ("capacity", C.c_uint32), ("count", C.c_uint32)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _aligned(nbytes, align=64):
    raw = np.zeros(nbytes + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


def aligned_array(n, dtype, align=64):
    dtype = np.dtype(dtype)
    return _aligned(max(n, 1) * dtype.itemsize, align).view(dtype)[:n] if n else _aligned(dtype.itemsize, align).view(dtype)[:0]


def load(fast=False, gpu_dropin=False):
    """gpu_dropin=True loads the same shim linked against nudge_b200's namespace-nudge drop-in instead of nudge.cpp."""
    path = os.path.join(HERE, "_ref", "libnudge_gpu_shim.so" if gpu_dropin else ("libnudge_ref_fast.so" if fast else "libnudge_ref.so"))
    if not os.path.exists(path):
        raise FileNotFoundError(path + " missing: run `make -C oracle ref` in the build container")
    lib = C.CDLL(path)
    lib.ref_read_cached_impulses.restype = C.c_void_p
    lib.ref_setup_contact_constraints.restype = C.c_void_p
    lib.ref_constraints_batches.restype = C.c_uint
    lib.ref_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
    lib.ref_step.argtypes = [C.c_void_p] * 7 + [C.c_size_t, C.c_float, C.c_uint, C.c_float, C.c_float, C.c_void_p]
    lib.ref_collide.argtypes = [C.c_void_p] * 6 + [C.c_size_t]
    for f in (lib.ref_write_cached_impulses, lib.ref_apply_impulses, lib.ref_update_cached_impulses):
        f.argtypes = [C.c_void_p] * (3 if f is lib.ref_write_cached_impulses else 2)
    lib.ref_read_cached_impulses.argtypes = [C.c_void_p] * 3
    lib.ref_setup_contact_constraints.argtypes = [C.c_void_p] * 5
    lib.ref_impulses_get.argtypes = [C.c_void_p, C.c_uint] + [C.c_void_p] * 5 + [C.c_uint]
    lib.ref_constraints_get.argtypes = [C.c_void_p] * 6
    lib.ref_constraints_batches.argtypes = [C.c_void_p]
    lib.ref_rcp.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    lib.ref_rsqrt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint]
    return lib


class RefSim:
    """Runs the unmodified reference on a nudge_b200.scenes.Scene (must fit the reference's uint16 limits)."""

    def __init__(self, scene, fast=False, ftz=False, contact_capacity=None, arena_mb=256, gpu_dropin=False):
        from nudge_b200 import scenes as S
        assert scene.fits_reference(), "scene exceeds the reference's limits (nudge.cpp:3010, nudge.h:68-71)"
        self.S = S
        self.lib = load(fast, gpu_dropin)
        self.lib.ref_set_ftz_daz(1 if ftz else 0)
        self.scene = scene
        nb = scene.n_bodies
        cap = contact_capacity or max(1024, nb * 64)
        self.cap = cap
        A = aligned_array
        self.transforms = A(nb, S.TRANSFORM); self.transforms[:] = scene.transforms
        self.properties = A(nb, S.PROPERTIES); self.properties[:] = scene.properties
        self.momentum = A(nb, S.MOMENTUM); self.momentum[:] = scene.momentum
        self.idle = A(nb, np.uint8); self.idle[:] = scene.idle
        self.box_tags = A(scene.n_boxes, np.uint16); self.box_tags[:] = scene.box_tags
        self.box_data = A(scene.n_boxes, S.BOX); self.box_data[:] = scene.box_data
        self.box_transforms = A(scene.n_boxes, S.TRANSFORM); self.box_transforms[:] = scene.box_transforms
        self.sphere_tags = A(scene.n_spheres, np.uint16); self.sphere_tags[:] = scene.sphere_tags
        self.sphere_data = A(scene.n_spheres, S.SPHERE); self.sphere_data[:] = scene.sphere_data
        self.sphere_transforms = A(scene.n_spheres, S.TRANSFORM); self.sphere_transforms[:] = scene.sphere_transforms
        self.connections = A(len(scene.connections), S.PAIR16)
        self.connections["a"] = scene.connections["a"]; self.connections["b"] = scene.connections["b"]
        self.contact_data = A(cap, S.CONTACT)
        self.contact_bodies = A(cap, S.PAIR16)
        self.contact_tags = A(cap, np.uint64)
        self.sleeping_pairs = A(cap, np.uint32)
        self.active_indices = A(nb, np.uint16)
        self.cache_tags = A(cap, np.uint64)
        self.cache_data = A(cap, S.IMPULSE)
        self.arena_buf = _aligned(arena_mb << 20, 4096)

        self.bodies = BodyData(_ptr(self.transforms), _ptr(self.properties), _ptr(self.momentum), _ptr(self.idle), nb)
        self.colliders = ColliderData(_Shapes(_ptr(self.box_tags), _ptr(self.box_data), _ptr(self.box_transforms), scene.n_boxes),
                                      _Shapes(_ptr(self.sphere_tags), _ptr(self.sphere_data), _ptr(self.sphere_transforms), scene.n_spheres))
        self.conn = BodyConnections(_ptr(self.connections), len(self.connections))
        self.contacts = ContactData(_ptr(self.contact_data), _ptr(self.contact_bodies), _ptr(self.contact_tags), cap, 0, _ptr(self.sleeping_pairs), 0)
        self.active = ActiveBodies(_ptr(self.active_indices), nb, 0)
        self.cache = ContactCache(_ptr(self.cache_tags), _ptr(self.cache_data), cap, 0)
        self.arena = Arena(self.arena_buf.ctypes.data, self.arena_buf.size)
        self.impulses = None
        self.constraints = None

    # ---- the seven API calls (nudge.h:134-146) ----
    def collide(self):
        self.arena = Arena(self.arena_buf.ctypes.data, self.arena_buf.size)  # example/main.cpp:282
        self.lib.ref_collide(C.byref(self.active), C.byref(self.contacts), C.byref(self.bodies), C.byref(self.colliders), C.byref(self.conn),
                             self.arena.data, self.arena.size)

    def apply_gravity_damping(self):
        """example/main.cpp:291-305 in float32."""
        f = np.float32
        dt = f(self.scene.time_step)
        damping = f(f(1.0) - dt * f(self.scene.damping))
        idx = self.active_indices[:self.active.count].astype(np.int64)
        m = self.momentum
        vy = m["velocity"][idx, 1] - f(f(self.scene.gravity) * dt)
        m["velocity"][idx, 1] = vy
        m["velocity"][idx] = m["velocity"][idx] * damping
        m["angular_velocity"][idx] = m["angular_velocity"][idx] * damping

    def read_cached_impulses(self):
        self.impulses = self.lib.ref_read_cached_impulses(C.byref(self.cache), C.byref(self.contacts), C.byref(self.arena))

    def setup_contact_constraints(self):
        self.constraints = self.lib.ref_setup_contact_constraints(C.byref(self.active), C.byref(self.contacts), C.byref(self.bodies), self.impulses, C.byref(self.arena))

    def apply_impulses(self):
        self.lib.ref_apply_impulses(self.constraints, C.byref(self.bodies))

    def update_cached_impulses(self):
        self.lib.ref_update_cached_impulses(self.constraints, self.impulses)

    def write_cached_impulses(self):
        self.lib.ref_write_cached_impulses(C.byref(self.cache), C.byref(self.contacts), self.impulses)

    def advance(self):
        self.lib.ref_advance(C.byref(self.active), C.byref(self.bodies), float(self.scene.time_step))

    def step(self, phases=None):
        """One sub-step of example/main.cpp:274-328 entirely inside the shim (used for timing)."""
        self.lib.ref_step(C.byref(self.active), C.byref(self.contacts), C.byref(self.bodies), C.byref(self.colliders), C.byref(self.conn),
                          C.byref(self.cache), self.arena_buf.ctypes.data, self.arena_buf.size, float(self.scene.time_step),
                          int(self.scene.iterations), float(self.scene.gravity), float(self.scene.damping),
                          phases.ctypes.data if phases is not None else None)

    def step_staged(self):
        """Same step through the seven calls (bit-identical to step(); checked in tests)."""
        self.collide()
        self.apply_gravity_damping()
        self.read_cached_impulses()
        self.setup_contact_constraints()
        for _ in range(int(self.scene.iterations)):
            self.apply_impulses()
        self.update_cached_impulses()
        self.write_cached_impulses()
        self.advance()

    # ---- views of results ----
    def contacts_view(self):
        n = self.contacts.count
        return dict(count=n, data=self.contact_data[:n].copy(), bodies=self.contact_bodies[:n].copy(), tags=self.contact_tags[:n].copy(),
                    sleeping=self.sleeping_pairs[:self.contacts.sleeping_count].copy(), active=self.active_indices[:self.active.count].copy())

    def impulses_view(self):
        n = self.contacts.count
        sorted_contacts = np.zeros(n, np.uint32)
        data = np.zeros(n, self.S.IMPULSE)
        cc = C.c_uint(0)
        cap = self.cache.count
        ctags = np.zeros(cap, np.uint64)
        cdata = np.zeros(cap, self.S.IMPULSE)
        self.lib.ref_impulses_get(self.impulses, n, _ptr(sorted_contacts), _ptr(data), C.byref(cc), _ptr(ctags), _ptr(cdata), cap)
        return dict(sorted=sorted_contacts, data=data, culled_tags=ctags[:cc.value], culled_data=cdata[:cc.value])

    def constraints_view(self):
        nb = self.lib.ref_constraints_batches(self.constraints)
        lanes = nb * 8
        c2c = np.zeros(lanes, np.uint32); a = np.zeros(lanes, np.uint32); b = np.zeros(lanes, np.uint32)
        rows = np.zeros((lanes, 39), np.float32); states = np.zeros((lanes, 3), np.float32)
        self.lib.ref_constraints_get(self.constraints, _ptr(c2c), _ptr(a), _ptr(b), _ptr(rows), _ptr(states))
        return dict(batches=nb, contact=c2c, a=a, b=b, rows=rows, states=states)

    def cache_view(self):
        n = self.cache.count
        return dict(tags=self.cache_tags[:n].copy(), data=self.cache_data[:n].copy())

    def export_state(self, scene=None):
        """Copies the current body state back into a Scene (to restart any implementation from identical state)."""
        s = (scene or self.scene).copy()
        s.transforms[:] = self.transforms; s.momentum[:] = self.momentum; s.idle[:] = self.idle
        return s
