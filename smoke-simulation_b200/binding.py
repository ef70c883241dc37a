"""ctypes binding of include/smoke_b200.h -- the host-side mirror of the reference interface.

``SmokeSim`` mirrors the reference's entry points (project/smokeSimulation.cuh:4-17) one to one:
``SmokeSim(W,H,D, smoke0)`` = initializeVolume, ``add_obstacle`` / ``add_source`` / ``update_object_pos``,
``gravity`` / ``buoyancy`` properties (the getGravity()/getBuoyancy() pointers), ``step(dt, out)`` = simulate,
``close()`` = deleteVolume.  There is no CPU fallback: if the CUDA library is missing or no device is
usable, construction raises ``SmokeError``.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "libsmoke_b200.so")
HEADER = os.path.join(ROOT, "include", "smoke_b200.h")

SMOKE, U, V, W, MASK = 0, 1, 2, 3, 4
NOW, PAST, BUF0, BUF1 = 0, 1, 2, 3
STAGES = ("fill", "force", "pressure", "advect_vel", "advect_smoke", "readback")

_f = C.c_float
_vp = C.c_void_p
_i = C.c_int


PASS_KERNELS = {"auto": 0, "reg": 1, "tma": 2}
DEFAULT_PASS_KERNEL = None   # tests set this to run every case with each pass kernel (SmokeSim.__init__ applies it)


class SmokeError(RuntimeError):
    pass


class HaloRegion(C.Structure):
    _fields_ = [("side", _i), ("send_ptr", _vp), ("recv_ptr", _vp), ("send_bytes", C.c_size_t), ("recv_bytes", C.c_size_t)]


EXCHANGE_FN = C.CFUNCTYPE(_i, _vp, _i, C.POINTER(HaloRegion), _i, _vp)

_lib = None


def build_library():
    """Compile libsmoke_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "csrc")], check=True)


def declared_symbols():
    """Every function include/smoke_b200.h declares (used by the ABI test)."""
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(smk_[a-z_0-9]+)\s*\(", txt)) - {"smk_exchange_fn"})


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SmokeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build() "
                         "(there is no CPU fallback for the smoke step)")
    L = C.CDLL(LIB_PATH)
    L.smk_create.argtypes = [C.POINTER(_vp), C.c_uint, C.c_uint, C.c_uint, _vp]
    L.smk_create_slab.argtypes = [C.POINTER(_vp)] + [C.c_uint] * 6 + [_vp]
    L.smk_slab_geometry.argtypes = [C.c_uint] * 6 + [C.POINTER(_i)]
    L.smk_slab_plan.argtypes = [C.c_uint] * 6 + [_i, _i, _i, C.POINTER(_i), _i]
    L.smk_slab_regions.argtypes = [C.c_uint] * 6 + [_i, C.POINTER(_i), _i]
    L.smk_pass_schedule.argtypes = [C.c_uint, C.c_uint, _i, _i, _i, _i, C.POINTER(_i), _i, C.POINTER(_i), _i, C.POINTER(_i)]
    L.smk_destroy.argtypes = [_vp]
    L.smk_add_obstacle.argtypes = [_vp] + [_f] * 7
    L.smk_add_source.argtypes = [_vp] + [_f] * 4
    L.smk_update_object_pos.argtypes = [_vp, _i, _f, _f, _f]
    L.smk_gravity_ptr.argtypes = [_vp]
    L.smk_gravity_ptr.restype = C.POINTER(_f)
    L.smk_buoyancy_ptr.argtypes = [_vp]
    L.smk_buoyancy_ptr.restype = C.POINTER(_f)
    L.smk_set_solver.argtypes = [_vp, _i, _i, _i]
    L.smk_set_pass_kernel.argtypes = [_vp, _i]
    L.smk_last_pass_kernel.argtypes = [_vp]
    L.smk_set_pass_ctas.argtypes = [_vp, _i]
    L.smk_set_readback_box.argtypes = [_vp, _i]
    L.smk_last_pass_ctas.argtypes = [_vp]
    L.smk_set_obstacle_mode.argtypes = [_vp, _i]
    L.smk_read_density_half.argtypes = [_vp, _vp]
    L.smk_step.argtypes = [_vp, _f, _vp]
    L.smk_step_async.argtypes = [_vp, _f, _vp]
    L.smk_sync.argtypes = [_vp]
    L.smk_register_host.argtypes = [_vp, _vp, C.c_size_t]
    L.smk_unregister_host.argtypes = [_vp, _vp]
    L.smk_hash_owned.argtypes = [_vp, C.POINTER(C.c_ulonglong)]
    L.smk_hash_range.argtypes = [_vp, _i, _i, _i, _i, C.POINTER(C.c_ulonglong)]
    L.smk_density_device.argtypes = [_vp]
    L.smk_density_device.restype = _vp
    L.smk_set_stream.argtypes = [_vp, _vp]
    L.smk_copy_density_to_array.argtypes = [_vp, _vp]
    L.smk_bind_density_array.argtypes = [_vp, _vp]
    L.smk_set_mask_bits.argtypes = [_vp, _vp]
    L.smk_get_mask_bits.argtypes = [_vp, _vp]
    L.smk_test_array_create.argtypes = [C.c_uint] * 3
    L.smk_test_array_create.restype = _vp
    L.smk_test_array_read.argtypes = [_vp, _vp] + [C.c_uint] * 3
    L.smk_test_array_destroy.argtypes = [_vp]
    L.smk_stage_flip.argtypes = [_vp]
    L.smk_stage_fill.argtypes = [_vp]
    L.smk_stage_force_clamp.argtypes = [_vp, _f]
    L.smk_stage_pressure_halfsweep.argtypes = [_vp, _i]
    L.smk_stage_pressure.argtypes = [_vp]
    L.smk_stage_advect_velocity.argtypes = [_vp, _f]
    L.smk_stage_advect_smoke.argtypes = [_vp, _f]
    L.smk_get_field.argtypes = [_vp, _i, _i, _vp]
    L.smk_set_field.argtypes = [_vp, _i, _i, _vp]
    L.smk_index_now.argtypes = [_vp]
    L.smk_max_divergence.argtypes = [_vp, C.POINTER(_f)]
    L.smk_stage_time.argtypes = [_vp, _i, C.POINTER(C.c_double), C.POINTER(C.c_long)]
    L.smk_reset_timers.argtypes = [_vp]
    L.smk_launch_count.argtypes = [_vp]
    L.smk_launch_count.restype = C.c_long
    L.smk_readback_bytes.argtypes = [_vp]
    L.smk_readback_bytes.restype = C.c_ulonglong
    L.smk_selfcheck_omega.argtypes = [_i, C.c_uint, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.smk_set_exchange.argtypes = [_vp, EXCHANGE_FN, _vp]
    L.smk_exec_op.argtypes = [_vp, C.POINTER(_i), _f]
    L.smk_p2p_export.argtypes = [_vp, C.c_char_p]
    L.smk_p2p_attach_ipc.argtypes = [_vp, _i, C.c_char_p]
    L.smk_p2p_arena.argtypes = [_vp]
    L.smk_p2p_arena.restype = _vp
    L.smk_p2p_attach_ptr.argtypes = [_vp, _i, _vp]
    L.smk_p2p_presignal.argtypes = [_vp]
    L.smk_exchange_count.argtypes = [_vp]
    L.smk_exchange_count.restype = C.c_long
    L.smk_slab_plan_p2p.argtypes = [C.c_uint] * 6 + [_i, _i, _i, C.POINTER(_i), _i]
    L.smk_last_error.argtypes = [_vp]
    L.smk_last_error.restype = C.c_char_p
    _lib = L
    return L


def selfcheck_omega(first=0, count=1 << 32, device=0):
    """(mismatches, ties) of the binary32 over-relaxation product against the double-precision one, on the device."""
    L = load_library()
    bad, ties = C.c_ulonglong(0), C.c_ulonglong(0)
    rc = L.smk_selfcheck_omega(device, first, count, C.byref(bad), C.byref(ties))
    if rc != 0:
        raise SmokeError(f"smk_selfcheck_omega failed: {rc}")
    return int(bad.value), int(ties.value)


def field_shape(field, W_, H_, D_):
    if field in (SMOKE, MASK):
        return (D_, H_, W_)
    return (D_ + 1, H_ + 1, W_ + 1)


def field_dtype(field):
    return np.uint8 if field == MASK else np.float32


class SmokeSim:
    """One simulation volume on the current CUDA device (reference: the process globals of cu:16-58)."""

    def __init__(self, W_, H_, D_, smoke0=None, slab=None, ghost=8):
        """slab = (rank, world) creates one z-slab of a multi-GPU run (smk_create_slab)."""
        self.L = load_library()
        self.W, self.H, self.D = int(W_), int(H_), int(D_)
        self.h = _vp()
        p = None
        if smoke0 is not None:
            smoke0 = np.ascontiguousarray(smoke0, dtype=np.float32)
            assert smoke0.size == self.W * self.H * self.D
            p = smoke0.ctypes.data_as(_vp)
        if slab is None:
            rc = self.L.smk_create(C.byref(self.h), self.W, self.H, self.D, p)
        else:
            rc = self.L.smk_create_slab(C.byref(self.h), self.W, self.H, self.D, int(slab[0]), int(slab[1]), int(ghost), p)
        if rc != 0:
            raise SmokeError(f"smk_create failed ({rc}): {self.L.smk_last_error(None).decode()}")
        self._cb = None
        if DEFAULT_PASS_KERNEL is not None:
            self.set_pass_kernel(DEFAULT_PASS_KERNEL)

    # -- lifetime
    def close(self):
        if getattr(self, "h", None):
            self.L.smk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SmokeError(f"libsmoke_b200 error {rc}: {self.L.smk_last_error(self.h).decode()}")

    # -- scene / parameters (cu:88-109, 31-36)
    def add_obstacle(self, x, y, z, vx, vy, vz, r):
        i = self.L.smk_add_obstacle(self.h, x, y, z, vx, vy, vz, r)
        if i < 0:
            self._ck(-i)
        return i

    def add_source(self, x, y, z, r):
        i = self.L.smk_add_source(self.h, x, y, z, r)
        if i < 0:
            self._ck(-i)
        return i

    def update_object_pos(self, i, x, y, z): self._ck(self.L.smk_update_object_pos(self.h, i, x, y, z))

    @property
    def gravity(self): return self.L.smk_gravity_ptr(self.h)[0]
    @gravity.setter
    def gravity(self, v): self.L.smk_gravity_ptr(self.h)[0] = v
    @property
    def buoyancy(self): return self.L.smk_buoyancy_ptr(self.h)[0]
    @buoyancy.setter
    def buoyancy(self, v): self.L.smk_buoyancy_ptr(self.h)[0] = v

    def set_params(self, g, a):
        self.gravity, self.buoyancy = g, a

    def read_density_half(self, out=None):
        """Density of the last step as float16 (opt-in extension N4); `out` = (D, H, W) float16 array or None."""
        if out is None:
            out = np.zeros((self.D, self.H, self.W), dtype=np.float16)
        self._ck(self.L.smk_read_density_half(self.h, out.ctypes.data_as(_vp)))
        return out

    def set_obstacle_mode(self, union_mode): self._ck(self.L.smk_set_obstacle_mode(self.h, int(union_mode)))

    def set_pass_kernel(self, kind): self._ck(self.L.smk_set_pass_kernel(self.h, PASS_KERNELS.get(kind, kind)))
    def last_pass_kernel(self): return {1: "reg", 2: "tma"}.get(int(self.L.smk_last_pass_kernel(self.h)), "none")
    def set_solver(self, variant=0, iterations=30, fuse=0): self._ck(self.L.smk_set_solver(self.h, variant, iterations, fuse))
    def set_pass_ctas(self, nctas): self._ck(self.L.smk_set_pass_ctas(self.h, int(nctas)))
    def last_pass_ctas(self): return int(self.L.smk_last_pass_ctas(self.h))
    def set_readback_box(self, on): self._ck(self.L.smk_set_readback_box(self.h, int(on)))
    def set_stream(self, cuda_stream): self._ck(self.L.smk_set_stream(self.h, cuda_stream))

    # -- the step (cu:774-819)
    def step(self, dt, out=None):
        """simulate(): `out` = host array (W*H*D float32) that receives the density, or None."""
        p = None if out is None else out.ctypes.data_as(_vp)
        self._ck(self.L.smk_step(self.h, dt, p))

    def step_ptr(self, dt, host_ptr): self._ck(self.L.smk_step(self.h, dt, host_ptr))
    def step_async(self, dt, host_ptr=None): self._ck(self.L.smk_step_async(self.h, dt, host_ptr))
    def sync(self): self._ck(self.L.smk_sync(self.h))
    def density_device(self): return self.L.smk_density_device(self.h)
    def register_host(self, arr): self._ck(self.L.smk_register_host(self.h, arr.ctypes.data_as(_vp), arr.nbytes))
    def unregister_host(self, arr): self._ck(self.L.smk_unregister_host(self.h, arr.ctypes.data_as(_vp)))
    def copy_density_to_array(self, cuda_array): self._ck(self.L.smk_copy_density_to_array(self.h, cuda_array))
    def bind_density_array(self, cuda_array): self._ck(self.L.smk_bind_density_array(self.h, cuda_array))

    def set_mask_bits(self, bits):
        """bits: uint8 array of (W*H*D + 7) // 8 bytes, np.packbits(mask.ravel(), bitorder='little')."""
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        assert b.size == (self.W * self.H * self.D + 7) // 8
        self._ck(self.L.smk_set_mask_bits(self.h, b.ctypes.data_as(_vp)))

    def get_mask_bits(self):
        b = np.zeros((self.W * self.H * self.D + 7) // 8, dtype=np.uint8)
        self._ck(self.L.smk_get_mask_bits(self.h, b.ctypes.data_as(_vp)))
        return b

    # -- stages
    def flip(self): self._ck(self.L.smk_stage_flip(self.h))
    def fill(self): self._ck(self.L.smk_stage_fill(self.h))
    def force_clamp(self, dt): self._ck(self.L.smk_stage_force_clamp(self.h, dt))
    def integrate(self, dt): raise NotImplementedError("forcing and clamp are one fused kernel: use force_clamp")
    def pressure_halfsweep(self, offset): self._ck(self.L.smk_stage_pressure_halfsweep(self.h, offset))
    def pressure(self): self._ck(self.L.smk_stage_pressure(self.h))
    def advect_velocity(self, dt): self._ck(self.L.smk_stage_advect_velocity(self.h, dt))
    def advect_smoke(self, dt): self._ck(self.L.smk_stage_advect_smoke(self.h, dt))
    def index_now(self): return self.L.smk_index_now(self.h)

    # -- fields
    def get_field(self, field, which=NOW):
        a = np.zeros(field_shape(field, self.W, self.H, self.D), dtype=field_dtype(field))
        self._ck(self.L.smk_get_field(self.h, field, which, a.ctypes.data_as(_vp)))
        return a

    def set_field(self, field, which, arr):
        a = np.ascontiguousarray(arr, dtype=field_dtype(field))
        assert a.shape == field_shape(field, self.W, self.H, self.D), (a.shape, field)
        self._ck(self.L.smk_set_field(self.h, field, which, a.ctypes.data_as(_vp)))

    def max_divergence(self):
        v = _f()
        self._ck(self.L.smk_max_divergence(self.h, C.byref(v)))
        return float(v.value)

    HASH_NAMES = ("u_now", "v_now", "w_now", "u_past", "v_past", "w_past", "density_past")

    def hash_owned(self):
        """64-bit content hashes of the owned planes, [u, v, w now, u, v, w past, density past] (device-side)."""
        out = (C.c_ulonglong * 7)()
        self._ck(self.L.smk_hash_owned(self.h, out))
        return [int(v) for v in out]

    def hash_range(self, node_lo, node_hi, cell_lo, cell_hi):
        out = (C.c_ulonglong * 7)()
        self._ck(self.L.smk_hash_range(self.h, node_lo, node_hi, cell_lo, cell_hi, out))
        return [int(v) for v in out]

    # -- measurement
    def stage_times(self):
        out = {}
        for i, n in enumerate(STAGES):
            ms, k = C.c_double(), C.c_long()
            self._ck(self.L.smk_stage_time(self.h, i, C.byref(ms), C.byref(k)))
            out[n] = (ms.value, k.value)
        return out

    def reset_timers(self): self._ck(self.L.smk_reset_timers(self.h))
    def launch_count(self): return int(self.L.smk_launch_count(self.h))
    def readback_bytes(self): return int(self.L.smk_readback_bytes(self.h))

    def exec_op(self, op, dt):
        """Run one plan op; `op` = (name-or-kind, a, b, p0, p1) as returned by slab.plan()."""
        from .slab import OP_NAMES
        kind = OP_NAMES.index(op[0]) if isinstance(op[0], str) else int(op[0])
        arr = (_i * 5)(kind, *[int(v) for v in op[1:5]])
        self._ck(self.L.smk_exec_op(self.h, arr, dt))

    # -- peer-memory halo path
    def p2p_export(self):
        """64-byte CUDA IPC handle of this slab's arena (send it to the neighbour processes)."""
        buf = C.create_string_buffer(64)
        self._ck(self.L.smk_p2p_export(self.h, buf))
        return buf.raw

    def p2p_attach_ipc(self, side, handle): self._ck(self.L.smk_p2p_attach_ipc(self.h, side, bytes(handle)))
    def p2p_arena(self): return self.L.smk_p2p_arena(self.h)
    def p2p_attach_ptr(self, side, ptr): self._ck(self.L.smk_p2p_attach_ptr(self.h, side, ptr))
    def exchange_count(self): return int(self.L.smk_exchange_count(self.h))
    def p2p_presignal(self): self._ck(self.L.smk_p2p_presignal(self.h))

    def set_exchange(self, pyfunc):
        """pyfunc(set_id, [(side, send_ptr, recv_ptr, send_bytes, recv_bytes), ...], stream_ptr) -> int"""
        def tramp(ctx, set_id, regions, n, stream):
            try:
                return int(pyfunc(set_id, [(regions[i].side, regions[i].send_ptr, regions[i].recv_ptr, regions[i].send_bytes,
                                            regions[i].recv_bytes) for i in range(n)], stream) or 0)
            except Exception:  # pragma: no cover - surfaced as SMK_ERR_TRANSPORT
                import traceback
                traceback.print_exc()
                return 5
        self._cb = EXCHANGE_FN(tramp)
        self._ck(self.L.smk_set_exchange(self.h, self._cb, None))
