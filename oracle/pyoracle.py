"""ctypes front-ends for the parity oracles.  TEST INFRASTRUCTURE ONLY.

Three engines with one interface (create / add_* / step / stage calls / get_field / set_field):

* ``Oracle``  -- oracle/liboracle.so, my C restatement of the reference step (smoke_oracle.c).
* ``RefCPU``  -- oracle/_ref/libref_cpu.so, the reference's own kernel bodies looped on the host.
* ``RefGPU``  -- oracle/_ref/libref_gpu.so, the reference step itself built headless for sm_100a
                 (process-global singleton, exactly like the reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_ORACLE = os.path.join(HERE, "liboracle.so")
LIB_REF_CPU = os.path.join(HERE, "_ref", "libref_cpu.so")
LIB_REF_GPU = os.path.join(HERE, "_ref", "libref_gpu.so")

SMOKE, U, V, W, MASK = 0, 1, 2, 3, 4
NOW, PAST, BUF0, BUF1 = 0, 1, 2, 3

_f = C.c_float
_vp = C.c_void_p


def build(ref=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref/*."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


def have_ref_cpu():
    return os.path.exists(LIB_REF_CPU)


def have_ref_gpu():
    return os.path.exists(LIB_REF_GPU)


def field_shape(field, W_, H_, D_):
    """numpy shape (z, y, x) of a field in the reference layout (x fastest)."""
    if field in (SMOKE, MASK):
        return (D_, H_, W_)
    return (D_ + 1, H_ + 1, W_ + 1)


def field_dtype(field):
    return np.uint8 if field == MASK else np.float32


class _Engine:
    """Shared python-side helpers; subclasses bind the C symbols."""

    def __init__(self, W_, H_, D_):
        self.W, self.H, self.D = int(W_), int(H_), int(D_)

    def get_field(self, field, which=NOW):
        a = np.empty(field_shape(field, self.W, self.H, self.D), dtype=field_dtype(field))
        self._get(field, which, a.ctypes.data_as(_vp))
        return a

    def set_field(self, field, which, arr):
        a = np.ascontiguousarray(arr, dtype=field_dtype(field))
        assert a.shape == field_shape(field, self.W, self.H, self.D), (a.shape, field)
        self._set(field, which, a.ctypes.data_as(_vp))

    def max_divergence(self):
        """max |div| over interior fluid cells of the 'now' velocities (formula cu:379-381)."""
        u, v, w, s = (self.get_field(f, NOW) for f in (U, V, W, MASK))
        return max_divergence(u, v, w, s)


def max_divergence(u, v, w, s):
    D_, H_, W_ = s.shape
    c = (slice(1, D_ - 1), slice(1, H_ - 1), slice(1, W_ - 1))
    u0 = u[1:D_ - 1, 1:H_ - 1, 1:W_ - 1]; u1 = u[1:D_ - 1, 1:H_ - 1, 2:W_]
    v0 = v[1:D_ - 1, 1:H_ - 1, 1:W_ - 1]; v1 = v[1:D_ - 1, 2:H_, 1:W_ - 1]
    w0 = w[1:D_ - 1, 1:H_ - 1, 1:W_ - 1]; w1 = w[2:D_, 1:H_ - 1, 1:W_ - 1]
    div = ((((-u0 + u1) + -v0) + v1) + -w0) + w1
    div = np.abs(div) * (s[c] != 0)
    return float(div.max()) if div.size else 0.0


class Oracle(_Engine):
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(LIB_ORACLE):
                build(ref=False)
            L = C.CDLL(LIB_ORACLE)
            L.orc_create.restype = _vp
            L.orc_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, _vp, C.c_int]
            L.orc_destroy.argtypes = [_vp]
            L.orc_set_threads.argtypes = [C.c_int]
            L.orc_get_threads.restype = C.c_int
            L.orc_add_obstacle.argtypes = [_vp] + [_f] * 7
            L.orc_add_source.argtypes = [_vp] + [_f] * 4
            L.orc_update_object_pos.argtypes = [_vp, C.c_int, _f, _f, _f]
            L.orc_set_params.argtypes = [_vp, _f, _f]
            L.orc_set_iterations.argtypes = [_vp, C.c_int]
            L.orc_set_solver.argtypes = [_vp, C.c_int]
            L.orc_set_obstacle_mode.argtypes = [_vp, C.c_int]
            L.orc_jacobi_iteration.argtypes = [_vp, _vp]
            L.orc_step.argtypes = [_vp, _f]
            L.orc_flip.argtypes = [_vp]
            L.orc_fill.argtypes = [_vp]
            for n in ("integrate", "clamp", "advect_velocity", "advect_smoke"):
                getattr(L, "orc_" + n).argtypes = [_vp, _f]
            L.orc_pressure_halfsweep.argtypes = [_vp, C.c_int]
            for n in ("integrate_r", "clamp_r", "advect_velocity_r", "advect_smoke_r"):
                getattr(L, "orc_" + n).argtypes = [_vp, _f, C.c_int, C.c_int]
            L.orc_pressure_halfsweep_r.argtypes = [_vp, C.c_int, C.c_int, C.c_int]
            L.orc_get_field.argtypes = [_vp, C.c_int, C.c_int, _vp]
            L.orc_set_field.argtypes = [_vp, C.c_int, C.c_int, _vp]
            L.orc_index_now.argtypes = [_vp]
            L.orc_max_divergence.argtypes = [_vp]
            L.orc_max_divergence.restype = _f
            cls._lib = L
        return cls._lib

    def __init__(self, W_, H_, D_, smoke0=None, contract=1):
        super().__init__(W_, H_, D_)
        L = self.lib()
        p = None
        if smoke0 is not None:
            smoke0 = np.ascontiguousarray(smoke0, dtype=np.float32)
            p = smoke0.ctypes.data_as(_vp)
        self.h = L.orc_create(self.W, self.H, self.D, p, int(contract))
        self.contract = contract

    def close(self):
        if self.h:
            self.lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def set_threads(cls, n): cls.lib().orc_set_threads(n)
    @classmethod
    def get_threads(cls): return cls.lib().orc_get_threads()
    def add_obstacle(self, x, y, z, vx, vy, vz, r): return self.lib().orc_add_obstacle(self.h, x, y, z, vx, vy, vz, r)
    def add_source(self, x, y, z, r): return self.lib().orc_add_source(self.h, x, y, z, r)
    def update_object_pos(self, i, x, y, z): self.lib().orc_update_object_pos(self.h, i, x, y, z)
    def set_params(self, g, a): self.lib().orc_set_params(self.h, g, a)
    def set_iterations(self, n): self.lib().orc_set_iterations(self.h, n)
    def set_solver(self, variant=0, iterations=30, fuse=0):
        self.lib().orc_set_solver(self.h, variant); self.lib().orc_set_iterations(self.h, iterations)

    def set_obstacle_mode(self, union_mode): self.lib().orc_set_obstacle_mode(self.h, int(union_mode))

    def jacobi_iteration(self):
        p = np.zeros(self.W * self.H * self.D, dtype=np.float32)
        self.lib().orc_jacobi_iteration(self.h, p.ctypes.data_as(_vp))

    def step(self, dt): self.lib().orc_step(self.h, dt)
    def flip(self): self.lib().orc_flip(self.h)
    def fill(self): self.lib().orc_fill(self.h)
    def integrate(self, dt): self.lib().orc_integrate(self.h, dt)
    def clamp(self, dt): self.lib().orc_clamp(self.h, dt)
    def pressure_halfsweep(self, offset): self.lib().orc_pressure_halfsweep(self.h, offset)
    def advect_velocity(self, dt): self.lib().orc_advect_velocity(self.h, dt)
    def advect_smoke(self, dt): self.lib().orc_advect_smoke(self.h, dt)
    def index_now(self): return self.lib().orc_index_now(self.h)
    def _get(self, f, w, p): self.lib().orc_get_field(self.h, f, w, p)
    def _set(self, f, w, p): self.lib().orc_set_field(self.h, f, w, p)
    # plane-range variants (z-slab emulation)
    def integrate_r(self, dt, za, zb): self.lib().orc_integrate_r(self.h, dt, za, zb)
    def clamp_r(self, dt, za, zb): self.lib().orc_clamp_r(self.h, dt, za, zb)
    def pressure_halfsweep_r(self, offset, za, zb): self.lib().orc_pressure_halfsweep_r(self.h, offset, za, zb)
    def advect_velocity_r(self, dt, za, zb): self.lib().orc_advect_velocity_r(self.h, dt, za, zb)
    def advect_smoke_r(self, dt, za, zb): self.lib().orc_advect_smoke_r(self.h, dt, za, zb)


class RefCPU(_Engine):
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(LIB_REF_CPU)
            L.refcpu_create.restype = _vp
            L.refcpu_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, _vp]
            L.refcpu_destroy.argtypes = [_vp]
            L.refcpu_set_threads.argtypes = [C.c_int]
            L.refcpu_get_threads.restype = C.c_int
            L.refcpu_add_obstacle.argtypes = [_vp] + [_f] * 7
            L.refcpu_add_source.argtypes = [_vp] + [_f] * 4
            L.refcpu_update_object_pos.argtypes = [_vp, C.c_int, _f, _f, _f]
            L.refcpu_set_params.argtypes = [_vp, _f, _f]
            L.refcpu_set_iterations.argtypes = [_vp, C.c_int]
            L.refcpu_step.argtypes = [_vp, _f]
            L.refcpu_flip.argtypes = [_vp]
            L.refcpu_fill.argtypes = [_vp]
            for n in ("integrate", "clamp", "advect_velocity", "advect_smoke"):
                getattr(L, "refcpu_" + n).argtypes = [_vp, _f]
            L.refcpu_pressure_halfsweep.argtypes = [_vp, C.c_int]
            L.refcpu_get_field.argtypes = [_vp, C.c_int, C.c_int, _vp]
            L.refcpu_set_field.argtypes = [_vp, C.c_int, C.c_int, _vp]
            L.refcpu_index_now.argtypes = [_vp]
            cls._lib = L
        return cls._lib

    def __init__(self, W_, H_, D_, smoke0=None):
        super().__init__(W_, H_, D_)
        p = None
        if smoke0 is not None:
            smoke0 = np.ascontiguousarray(smoke0, dtype=np.float32)
            p = smoke0.ctypes.data_as(_vp)
        self.h = self.lib().refcpu_create(self.W, self.H, self.D, p)

    def close(self):
        if self.h:
            self.lib().refcpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def set_threads(cls, n): cls.lib().refcpu_set_threads(n)
    @classmethod
    def get_threads(cls): return cls.lib().refcpu_get_threads()
    def add_obstacle(self, x, y, z, vx, vy, vz, r): return self.lib().refcpu_add_obstacle(self.h, x, y, z, vx, vy, vz, r)
    def add_source(self, x, y, z, r): return self.lib().refcpu_add_source(self.h, x, y, z, r)
    def update_object_pos(self, i, x, y, z): self.lib().refcpu_update_object_pos(self.h, i, x, y, z)
    def set_params(self, g, a): self.lib().refcpu_set_params(self.h, g, a)
    def set_iterations(self, n): self.lib().refcpu_set_iterations(self.h, n)
    def step(self, dt): self.lib().refcpu_step(self.h, dt)
    def flip(self): self.lib().refcpu_flip(self.h)
    def fill(self): self.lib().refcpu_fill(self.h)
    def integrate(self, dt): self.lib().refcpu_integrate(self.h, dt)
    def clamp(self, dt): self.lib().refcpu_clamp(self.h, dt)
    def pressure_halfsweep(self, offset): self.lib().refcpu_pressure_halfsweep(self.h, offset)
    def advect_velocity(self, dt): self.lib().refcpu_advect_velocity(self.h, dt)
    def advect_smoke(self, dt): self.lib().refcpu_advect_smoke(self.h, dt)
    def index_now(self): return self.lib().refcpu_index_now(self.h)
    def _get(self, f, w, p): self.lib().refcpu_get_field(self.h, f, w, p)
    def _set(self, f, w, p): self.lib().refcpu_set_field(self.h, f, w, p)


class RefGPU(_Engine):
    """The reference step on the GPU.  Process-global state (one instance alive at a time)."""
    _lib = None
    _alive = False

    @classmethod
    def lib(cls):
        if cls._lib is None:
            L = C.CDLL(LIB_REF_GPU)
            L.refgpu_init.argtypes = [_vp, C.c_uint, C.c_uint, C.c_uint]
            L.refgpu_add_obstacle.argtypes = [_f] * 7
            L.refgpu_add_source.argtypes = [_f] * 4
            L.refgpu_update_object_pos.argtypes = [C.c_int, _f, _f, _f]
            L.refgpu_set_params.argtypes = [_f, _f]
            L.refgpu_simulate.argtypes = [_vp, _f]
            for n in ("integrate", "clamp", "advect_velocity", "advect_smoke"):
                getattr(L, "refgpu_" + n).argtypes = [_f]
            L.refgpu_pressure_halfsweep.argtypes = [C.c_int]
            L.refgpu_get_field.argtypes = [C.c_int, C.c_int, _vp]
            L.refgpu_set_field.argtypes = [C.c_int, C.c_int, _vp]
            L.refgpu_time_kernels.argtypes = [_f, C.c_int]
            L.refgpu_time_kernels.restype = _f
            cls._lib = L
        return cls._lib

    def __init__(self, W_, H_, D_, smoke0=None):
        super().__init__(W_, H_, D_)
        assert not RefGPU._alive, "the reference keeps its state in process globals: one scene at a time"
        if smoke0 is None:
            smoke0 = np.zeros((self.D, self.H, self.W), dtype=np.float32)
        smoke0 = np.ascontiguousarray(smoke0, dtype=np.float32)
        self.host = np.zeros((self.D, self.H, self.W), dtype=np.float32)
        rc = self.lib().refgpu_init(smoke0.ctypes.data_as(_vp), self.W, self.H, self.D)
        assert rc == 0, rc
        RefGPU._alive = True

    def close(self):
        if RefGPU._alive:
            self.lib().refgpu_free()
            RefGPU._alive = False

    def add_obstacle(self, x, y, z, vx, vy, vz, r): return self.lib().refgpu_add_obstacle(x, y, z, vx, vy, vz, r)
    def add_source(self, x, y, z, r): return self.lib().refgpu_add_source(x, y, z, r)
    def update_object_pos(self, i, x, y, z): self.lib().refgpu_update_object_pos(i, x, y, z)
    def set_params(self, g, a): self.lib().refgpu_set_params(g, a)
    def step(self, dt): self.lib().refgpu_simulate(self.host.ctypes.data_as(_vp), dt)
    def flip(self): self.lib().refgpu_flip()
    def fill(self): self.lib().refgpu_fill()
    def integrate(self, dt): self.lib().refgpu_integrate(dt)
    def clamp(self, dt): self.lib().refgpu_clamp(dt)
    def pressure_halfsweep(self, offset): self.lib().refgpu_pressure_halfsweep(offset)
    def advect_velocity(self, dt): self.lib().refgpu_advect_velocity(dt)
    def advect_smoke(self, dt): self.lib().refgpu_advect_smoke(dt)
    def index_now(self): return self.lib().refgpu_index_now()
    def sync(self): return self.lib().refgpu_sync()
    def time_kernels(self, dt, ticks): return float(self.lib().refgpu_time_kernels(dt, ticks))

    def _get(self, f, w, p):
        rc = self.lib().refgpu_get_field(f, w, p)
        assert rc == 0, rc

    def _set(self, f, w, p):
        rc = self.lib().refgpu_set_field(f, w, p)
        assert rc == 0, rc


# ---- the deterministic scenes of SURVEY.md section 8(d): defined once, in the product package (the oracle may import
# the product, never the other way round)
import sys as _sys
_root = os.path.dirname(HERE)
if _root not in _sys.path:
    _sys.path.insert(0, _root)
from smoke_simulation_b200.scenes import SCENES, scaled_scene, setup_scene, tick_dt  # noqa: E402,F401
