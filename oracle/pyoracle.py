"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so (the widened CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import subprocess
import numpy as np
from nudge_b200 import abi, scenes as S

HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", HERE, "oracle"])
    lib = C.CDLL(path)
    V = C.c_void_p
    lib.nbo_collide.argtypes = [V] * 5
    lib.nbo_read_cached_impulses.argtypes = [V, V]; lib.nbo_read_cached_impulses.restype = V
    lib.nbo_write_cached_impulses.argtypes = [V, V, V]
    lib.nbo_setup_contact_constraints.argtypes = [V, V, V, V]; lib.nbo_setup_contact_constraints.restype = V
    lib.nbo_apply_impulses.argtypes = [V, V]
    lib.nbo_update_cached_impulses.argtypes = [V, V]
    lib.nbo_advance.argtypes = [V, V, C.c_float]
    lib.nbo_free_impulses.argtypes = [V]; lib.nbo_free_constraints.argtypes = [V]
    lib.nbo_last_pair_count.restype = C.c_uint32
    lib.nbo_last_pairs.argtypes = [V, V]; lib.nbo_last_morton_order.argtypes = [V]
    lib.nbo_impulses_get.argtypes = [V, C.c_uint32, V, V, V, V, V, V, C.c_uint32]
    lib.nbo_constraints_batches.argtypes = [V]; lib.nbo_constraints_batches.restype = C.c_uint32
    lib.nbo_constraints_get.argtypes = [V] * 6
    lib.nbo_rcp.argtypes = [V, V, C.c_uint32]; lib.nbo_rsqrt.argtypes = [V, V, C.c_uint32]
    return lib


class OracleSim(abi.HostState):
    """Runs the widened restatement on a Scene of any size."""

    def __init__(self, scene, ftz=False, contact_capacity=None):
        super().__init__(scene, contact_capacity)
        self.lib = load()
        self.lib.nbo_set_ftz_daz(1 if ftz else 0)
        self.impulses = None
        self.constraints = None

    def _free(self):
        if self.constraints: self.lib.nbo_free_constraints(self.constraints); self.constraints = None
        if self.impulses: self.lib.nbo_free_impulses(self.impulses); self.impulses = None

    def collide(self):
        self._free()
        self.contacts.capacity = self.cap
        self.lib.nbo_collide(C.byref(self.active), C.byref(self.contacts), C.byref(self.bodies), C.byref(self.colliders), C.byref(self.conn))
        if self.lib.nbo_last_overflow():
            raise RuntimeError("oracle: contact capacity %d exceeded" % self.cap)

    def read_cached_impulses(self):
        self.impulses = self.lib.nbo_read_cached_impulses(C.byref(self.cache), C.byref(self.contacts))

    def setup_contact_constraints(self):
        self.constraints = self.lib.nbo_setup_contact_constraints(C.byref(self.active), C.byref(self.contacts), C.byref(self.bodies), self.impulses)

    def apply_impulses(self):
        self.lib.nbo_apply_impulses(self.constraints, C.byref(self.bodies))

    def update_cached_impulses(self):
        self.lib.nbo_update_cached_impulses(self.constraints, self.impulses)

    def write_cached_impulses(self):
        self.lib.nbo_write_cached_impulses(C.byref(self.cache), C.byref(self.contacts), self.impulses)

    def advance(self):
        self.lib.nbo_advance(C.byref(self.active), C.byref(self.bodies), float(self.scene.time_step))

    def step(self):
        self.collide()
        self.apply_gravity_damping()
        self.read_cached_impulses()
        self.setup_contact_constraints()
        for _ in range(int(self.scene.iterations)):
            self.apply_impulses()
        self.update_cached_impulses()
        self.write_cached_impulses()
        self.advance()

    step_staged = step

    def pairs_view(self):
        n = self.lib.nbo_last_pair_count()
        lo = np.zeros(n, np.uint32); hi = np.zeros(n, np.uint32)
        self.lib.nbo_last_pairs(abi.ptr(lo), abi.ptr(hi))
        order = np.zeros(self.scene.n_colliders, np.uint32)
        self.lib.nbo_last_morton_order(abi.ptr(order))
        return dict(lo=lo, hi=hi, order=order)

    def impulses_view(self):
        n = self.contacts.count
        sorted_contacts = np.zeros(n, np.uint32); data = np.zeros(n, S.IMPULSE)
        cc = C.c_uint32(0); cap = max(self.cache.count, 1)
        ctags = np.zeros(cap, np.uint64); cfeat = np.zeros(cap, np.uint32); cdata = np.zeros(cap, S.IMPULSE)
        self.lib.nbo_impulses_get(self.impulses, n, abi.ptr(sorted_contacts), abi.ptr(data), C.byref(cc), abi.ptr(ctags), abi.ptr(cfeat), abi.ptr(cdata), cap)
        return dict(sorted=sorted_contacts, data=data, culled_tags=ctags[:cc.value], culled_features=cfeat[:cc.value], culled_data=cdata[:cc.value])

    def constraints_view(self):
        nb = self.lib.nbo_constraints_batches(self.constraints)
        lanes = nb * 8
        c2c = np.zeros(lanes, np.uint32); a = np.zeros(lanes, np.uint32); b = np.zeros(lanes, np.uint32)
        rows = np.zeros((lanes, 39), np.float32); states = np.zeros((lanes, 3), np.float32)
        self.lib.nbo_constraints_get(self.constraints, abi.ptr(c2c), abi.ptr(a), abi.ptr(b), abi.ptr(rows), abi.ptr(states))
        return dict(batches=nb, contact=c2c, a=a, b=b, rows=rows, states=states)
