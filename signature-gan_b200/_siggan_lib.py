"""ctypes binding of libsiggan.so (include/siggan.h) + the flat-buffer plumbing shared by the drop-in modules.

There is deliberately no fallback: if the shared library is missing or no CUDA device is present, every
compute entry point raises. PyTorch is used for device memory, streams and autograd bookkeeping only.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, List, Optional, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("SIGGAN_LIB", os.path.join(_HERE, "libsiggan.so"))

SG_PREC_BF16, SG_PREC_FP32 = 0, 1
SG_NET_G, SG_NET_D = 0, 1


class SgConfig(C.Structure):
    _fields_ = [("image_size", C.c_int), ("latent_dim", C.c_int), ("precision", C.c_int),
                ("leaky_slope", C.c_float), ("bn_eps", C.c_float), ("bn_momentum", C.c_float),
                ("g_act_slope", C.c_float), ("width_mult", C.c_int)]


class SgTrainState(C.Structure):
    _fields_ = [("g_params", C.c_void_p), ("g_running_stats", C.c_void_p), ("g_exp_avg", C.c_void_p),
                ("g_exp_avg_sq", C.c_void_p), ("d_params", C.c_void_p), ("d_exp_avg", C.c_void_p),
                ("d_exp_avg_sq", C.c_void_p), ("g_step", C.c_longlong), ("d_step", C.c_longlong),
                ("g_lr", C.c_float), ("d_lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float),
                ("eps", C.c_float), ("label_smoothing", C.c_float), ("dropout_p", C.c_float),
                ("seed", C.c_uint64), ("offset", C.c_uint64), ("masks_real", C.c_void_p),
                ("masks_fake", C.c_void_p), ("world_size", C.c_int), ("grad_scale", C.c_float)]


# name -> (restype, argtypes); must list every symbol include/siggan.h declares (tests check this).
_P, _I, _LL, _F, _U64, _SZ = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_uint64, C.c_size_t
SYMBOLS: Dict[str, Tuple[object, list]] = {
    "sg_abi_version": (_I, []),
    "sg_last_error": (C.c_char_p, []),
    "sg_launch_count": (C.c_ulonglong, []),
    "sg_profile_enable": (_I, [_P, _I]),
    "sg_profile_dump": (_LL, [_P, C.c_char_p, _SZ]),
    "sg_create": (_I, [C.POINTER(SgConfig), C.POINTER(_P)]),
    "sg_destroy": (None, [_P]),
    "sg_num_tensors": (_I, [_P, _I]),
    "sg_tensor_info": (_I, [_P, _I, _I, C.POINTER(C.c_char_p), C.POINTER(_LL), C.POINTER(_I * 4)]),
    "sg_param_count": (_LL, [_P, _I]),
    "sg_g_stat_count": (_LL, [_P]),
    "sg_g_num_bn": (_I, [_P]),
    "sg_g_bn_info": (_I, [_P, _I, C.POINTER(_LL), C.POINTER(_LL), C.POINTER(_I)]),
    "sg_g_workspace_bytes": (_SZ, [_P, _I]),
    "sg_d_workspace_bytes": (_SZ, [_P, _I]),
    "sg_d_mask_count": (_LL, [_P, _I]),
    "sg_d_grad_tail_offset": (_LL, [_P]),
    "sg_d_feature_count": (_LL, [_P]),
    "sg_g_forward": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "sg_g_backward": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "sg_d_forward": (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "sg_d_backward": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "sg_ws_offset": (_LL, [_P, _I, _I, _I, _I]),
    "sg_d_backward_layer": (_I, [_P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _P, _P]),
    "sg_g_backward_layer": (_I, [_P, _P, _P, _I, _P, _I, _I, _P, _P, _P]),
    "sg_dropout_masks": (_I, [_P, _U64, _U64, _I, _F, _P, _P]),
    "sg_bce_forward": (_I, [_P, _P, _I, _P, _P]),
    "sg_bce_backward": (_I, [_P, _P, _I, _P, _P, _P]),
    "sg_adam_step": (_I, [_P, _P, _P, _P, _LL, _F, _F, _F, _F, _LL, _P]),
    "sg_spectral_norm_weight": (_I, [_P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P]),
    "sg_spectral_norm_backward": (_I, [_P, _P, _P, _P, _I, _I, _P, _P, _P]),
    "sg_augment_params": (_I, [_P, _P, _I, _I, _P, _P]),
    "sg_augment_batch": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "sg_ink_stats": (_I, [_P, _I, _I, _F, _P, _P, _P, _P]),
    "sg_set_sync_batchnorm": (_I, [_P, _P, _P, _I, _P, _LL]),
    "sg_comm_nccl_version": (_I, []),
    "sg_comm_unique_id": (_I, [_P, _SZ]),
    "sg_comm_init": (_I, [_P, _P, _SZ, _I, _I]),
    "sg_comm_destroy": (_I, [_P]),
    "sg_comm_world_size": (_I, [_P]),
    "sg_allreduce_grads": (_I, [_P, _I, _P, _LL, _LL, _I, _P]),
    "sg_allreduce_join": (_I, [_P, _P]),
    "sg_g_grad_tail_offset": (_LL, [_P]),
    "sg_train_step": (_I, [_P, C.POINTER(SgTrainState), _P, _P, _P, _I, _P, _P, _P, _I, _P]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library() -> C.CDLL:
    """dlopen libsiggan.so and type every entry point. Raises if the library was not built."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise ImportError(
                    f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(or `make -C signature-gan_b200/csrc`). The B200 path has no CPU or eager fallback.")
            lib = C.CDLL(_LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load_library().sg_last_error()
        raise RuntimeError(f"siggan {what} failed: {msg.decode() if msg else rc}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def current_stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def precision_from_env() -> int:
    v = os.environ.get("SIGGAN_PRECISION", "bf16").lower()
    if v in ("bf16", "bfloat16", "0"):
        return SG_PREC_BF16
    if v in ("fp32", "float32", "validate", "1"):
        return SG_PREC_FP32
    raise ValueError(f"SIGGAN_PRECISION must be bf16 or fp32, got {v!r}")


class DropoutStream:
    """The ONE counter-based Dropout2d mask stream of this process (sg_dropout_masks / sg_train_step hash (seed, offset +
    index)): shared by the Discriminator module path and the fused training step, so the two never replay each other's
    masks; the seed is torch's initial seed mixed with the data-parallel rank, so replicas that share a seed still
    draw different masks; `state_dict()` goes into checkpoints so that a resumed run continues the stream."""

    def __init__(self) -> None:
        self.offset = 0

    @staticmethod
    def seed() -> int:
        rank = 0
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank = dist.get_rank()
        except Exception:
            rank = 0
        return (torch.initial_seed() ^ (rank * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF

    def advance(self, n: int) -> int:
        """Returns the offset of the next `n` mask values and moves past them."""
        o = self.offset
        self.offset += int(n)
        return o

    def state_dict(self) -> dict:
        return {"offset": int(self.offset)}

    def load_state_dict(self, sd: dict) -> None:
        self.offset = int(sd.get("offset", 0))


DROPOUT = DropoutStream()


class Context:
    """One sg_ctx per (device, image_size, latent_dim, precision), shared by G, D and the fused step."""
    _cache: Dict[tuple, "Context"] = {}

    def __init__(self, device: torch.device, image_size: int, latent_dim: int, precision: int,
                 leaky_slope: float = 0.2, bn_eps: float = 1e-5, bn_momentum: float = 0.1, g_act_slope: float = 0.0,
                 width_mult: int = 1):
        self.lib = load_library()
        self.device = device
        cfg = SgConfig(image_size, latent_dim, precision, leaky_slope, bn_eps, bn_momentum, g_act_slope, width_mult)
        handle = _P()
        with torch.cuda.device(device):
            check(self.lib.sg_create(C.byref(cfg), C.byref(handle)), "sg_create")
        self.handle = handle
        self.image_size, self.latent_dim, self.precision = image_size, latent_dim, precision

    @classmethod
    def get(cls, device: torch.device, image_size: int, latent_dim: int, precision: int, leaky_slope: float = 0.2,
            bn_eps: float = 1e-5, bn_momentum: float = 0.1, g_act_slope: float = 0.0, width_mult: int = 1) -> "Context":
        if device.type != "cuda":
            raise RuntimeError("siggan_b200 runs on CUDA (sm_100a) only; there is no CPU path "
                               f"(module is on {device}). Move the module with .to('cuda').")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        key = (device.index, image_size, latent_dim, precision, float(leaky_slope), float(bn_eps), float(bn_momentum),
               float(g_act_slope), int(width_mult))
        ctx = cls._cache.get(key)
        if ctx is None:
            ctx = cls(device, image_size, latent_dim, precision, leaky_slope, bn_eps, bn_momentum, g_act_slope, width_mult)
            cls._cache[key] = ctx
        return ctx

    def set_sync_batchnorm(self, group=None, enable: bool = True) -> None:
        """Synchronised BatchNorm for data-parallel runs (include/siggan.h: sg_set_sync_batchnorm): the Generator's
        training-mode BatchNorm sums are all-reduced over `group` (default: the world group) from inside the library
        through a ctypes callback, on the stream the library launches on. NCCL on the GPU box; gloo works too."""
        import torch.distributed as dist
        if not enable or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            check(self.lib.sg_set_sync_batchnorm(self.handle, None, None, 1, None, 0), "sg_set_sync_batchnorm")
            self._sync = None
            return
        buf = torch.empty(4 * 16 * 1024, dtype=torch.float32, device=self.device)   # fc BatchNorm1d: <= 16*512 channels
        base = buf.data_ptr()
        errors: list = []

        def _allreduce(user, ptr_, count, stream) -> int:
            try:
                off = (int(ptr_) - base) // 4
                if int(stream or 0) != torch.cuda.current_stream(self.device).cuda_stream:
                    raise RuntimeError("SyncBN all-reduce must run on the stream the step was launched on")
                dist.all_reduce(buf[off:off + int(count)], op=dist.ReduceOp.SUM, group=group)
                return 0
            except Exception as e:  # never let an exception cross the C frame
                errors.append(e)
                return -1

        cb = C.CFUNCTYPE(_I, _P, _P, _LL, _P)(_allreduce)
        check(self.lib.sg_set_sync_batchnorm(self.handle, C.cast(cb, _P), None, dist.get_world_size(group), ptr(buf),
                                             buf.numel()), "sg_set_sync_batchnorm")
        self._sync = (cb, buf, errors)   # keep the callback and its buffer alive as long as the library may call them

    def init_comm(self, group=None) -> int:
        """Library-owned NCCL communicator over the ranks of `group` (default: the world group) for this context's
        device (include/siggan.h: sg_comm_init). The 128-byte NCCL id travels from rank 0 through torch.distributed's
        host channel (any backend); afterwards the gradient all-reduces are issued by libsiggan itself
        (sg_allreduce_grads, and inside sg_train_step phase 0). Returns the world size (1 = nothing to do)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            return 1
        if self.comm_world() == world:
            return world
        ident = C.create_string_buffer(128)
        if rank == 0:
            check(self.lib.sg_comm_unique_id(ident, 128), "sg_comm_unique_id")
        box = [ident.raw]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        with torch.cuda.device(self.device):
            check(self.lib.sg_comm_init(self.handle, box[0], 128, rank, world), "sg_comm_init")
        return world

    def comm_world(self) -> int:
        return int(self.lib.sg_comm_world_size(self.handle))

    def allreduce_grads(self, net: int, flat_grad: torch.Tensor, offset: int = 0, count: int = -1, overlap: bool = False) -> None:
        check(self.lib.sg_allreduce_grads(self.handle, net, ptr(flat_grad), offset, count, 1 if overlap else 0,
                                          current_stream(self.device)), "sg_allreduce_grads")

    def allreduce_join(self) -> None:
        check(self.lib.sg_allreduce_join(self.handle, current_stream(self.device)), "sg_allreduce_join")

    def tensor_table(self, net: int) -> List[Tuple[str, int, Tuple[int, ...]]]:
        out = []
        for i in range(self.lib.sg_num_tensors(self.handle, net)):
            name, off, shape = C.c_char_p(), _LL(), (_I * 4)()
            check(self.lib.sg_tensor_info(self.handle, net, i, C.byref(name), C.byref(off), C.byref(shape)), "tensor_info")
            out.append((name.value.decode(), off.value, tuple(s for s in shape if s > 0)))
        return out

    def profile(self, on: bool) -> None:
        check(self.lib.sg_profile_enable(self.handle, 1 if on else 0), "sg_profile_enable")

    def profile_records(self):
        """[(name, ms, flops, bytes)] of every op recorded since profile(True); synchronises the device."""
        buf = C.create_string_buffer(1 << 20)
        n = self.lib.sg_profile_dump(self.handle, buf, len(buf))
        if n < 0:
            check(-1, "sg_profile_dump")
        out = []
        for line in buf.value.decode().splitlines():
            name, ms, fl, by = line.split("\t")
            out.append((name, float(ms), float(fl), float(by)))
        return out

    def param_count(self, net: int) -> int:
        return int(self.lib.sg_param_count(self.handle, net))

    def bn_table(self) -> List[Tuple[int, int, int]]:
        out = []
        for i in range(self.lib.sg_g_num_bn(self.handle)):
            m, v, ch = _LL(), _LL(), _I()
            check(self.lib.sg_g_bn_info(self.handle, i, C.byref(m), C.byref(v), C.byref(ch)), "bn_info")
            out.append((m.value, v.value, ch.value))
        return out


class FlatParams:
    """Keeps a module's parameters (and optionally BN running stats) as views into flat fp32 CUDA buffers, laid
    out exactly as the library's tensor table, and hands out gradient staging buffers that never alias a live
    `.grad` (so autograd's accumulate-or-steal logic stays correct)."""

    def __init__(self, module: torch.nn.Module, net: int):
        self.module = module
        self.net = net
        self.flat: Optional[torch.Tensor] = None
        self.stats: Optional[torch.Tensor] = None
        self._gbuf: List[Optional[torch.Tensor]] = [None, None]
        self.params: List[torch.nn.Parameter] = []
        self.layout: List[Tuple[int, int, Tuple[int, ...]]] = []
        self._names: List[str] = []

    def sync(self, ctx: Context, bns: Optional[list] = None) -> None:
        """(Re)build the flat buffers if any parameter no longer lives at its slot (first call, .to(), ...)."""
        table = ctx.tensor_table(self.net) if not self.layout else None
        if table is not None:
            # A spectral-normalised layer (torch.nn.utils.spectral_norm, disc…:61-62, 201-202) owns `weight_orig`
            # where the library's table says `weight`, and lists it after `bias`: bind by name, in table order.
            named = dict(self.module.named_parameters())
            self._names = [n if n in named else n + "_orig" for n in (t[0] for t in table)]
            if sorted(self._names) != sorted(named):
                raise RuntimeError(f"parameter table mismatch between module and libsiggan: {sorted(named)} vs {[t[0] for t in table]}")
        named = dict(self.module.named_parameters())
        params = [named[n] for n in self._names]
        if table is not None:
            for p, (_, off, shape) in zip(params, table):
                if tuple(p.shape) != shape:
                    raise RuntimeError(f"parameter shape mismatch: {tuple(p.shape)} vs {shape}")
            self.layout = [(off, int(torch.Size(shape).numel()), shape) for (_, off, shape) in table]
        self.params = params
        flat = self.flat
        ok = flat is not None and flat.device == params[0].device
        if ok:
            base = flat.data_ptr()
            for p, (off, n, _) in zip(params, self.layout):
                if p.data_ptr() != base + 4 * off or p.dtype != torch.float32:
                    ok = False
                    break
        if not ok:
            dev = params[0].device
            for p in params:
                if p.dtype != torch.float32:
                    raise RuntimeError("siggan_b200 keeps fp32 master parameters; .half()/.bfloat16() modules are unsupported")
            total = ctx.param_count(self.net)
            flat = torch.empty(total, dtype=torch.float32, device=dev)
            with torch.no_grad():
                for p, (off, n, shape) in zip(params, self.layout):
                    flat[off:off + n].copy_(p.data.reshape(-1))
                    p.data = flat[off:off + n].view(shape)
            self.flat = flat
            self._gbuf = [None, None]
        if bns is not None:
            tab = ctx.bn_table()
            st = self.stats
            good = st is not None and st.device == params[0].device
            if good:
                for bn, (mo, vo, ch) in zip(bns, tab):
                    if bn.running_mean.data_ptr() != st.data_ptr() + 4 * mo or bn.running_var.data_ptr() != st.data_ptr() + 4 * vo:
                        good = False
                        break
            if not good:
                total = int(ctx.lib.sg_g_stat_count(ctx.handle))
                st = torch.empty(total, dtype=torch.float32, device=params[0].device)
                with torch.no_grad():
                    for bn, (mo, vo, ch) in zip(bns, tab):
                        st[mo:mo + ch].copy_(bn.running_mean)
                        st[vo:vo + ch].copy_(bn.running_var)
                        bn._buffers["running_mean"] = st[mo:mo + ch]
                        bn._buffers["running_var"] = st[vo:vo + ch]
                self.stats = st

    def grad_staging(self) -> torch.Tensor:
        """A flat gradient buffer that no parameter's current .grad aliases."""
        busy = set()
        for k, buf in enumerate(self._gbuf):
            if buf is None:
                continue
            lo, hi = buf.data_ptr(), buf.data_ptr() + 4 * buf.numel()
            for p in self.params:
                g = p.grad
                if g is not None and lo <= g.data_ptr() < hi:
                    busy.add(k)
                    break
        for k in (0, 1):
            if k not in busy:
                if self._gbuf[k] is None:
                    self._gbuf[k] = torch.empty_like(self.flat)
                return self._gbuf[k]
        return torch.empty_like(self.flat)

    def grad_views(self, flat_grad: torch.Tensor) -> List[torch.Tensor]:
        return [flat_grad[off:off + n].view(shape) for (off, n, shape) in self.layout]

    def expose(self, flat_grad: torch.Tensor) -> None:
        """Publish a flat gradient buffer as the parameters' .grad (what autograd would have left behind)."""
        key = flat_grad.data_ptr()
        cache = getattr(self, "_view_cache", None)
        if cache is None or cache[0] != key:
            cache = (key, self.grad_views(flat_grad))
            self._view_cache = cache
        for p, v in zip(self.params, cache[1]):
            p.grad = v

    def fresh_grad(self) -> torch.Tensor:
        """A new flat gradient buffer for one autograd backward call. Autograd may keep (steal) views of it as
        `.grad`, or hold them in its input buffers while a second backward of the same network runs, so a buffer
        handed to autograd is never reused."""
        return torch.empty_like(self.flat)

    def flat_grad_if_contiguous(self) -> Optional[torch.Tensor]:
        """If every p.grad is the matching view of ONE flat buffer (what our backward hands to autograd), return a
        flat view over it (zero-copy); otherwise None."""
        g0 = self.params[0].grad
        if g0 is None or g0.dtype != torch.float32 or g0.device != self.flat.device:
            return None
        total = self.flat.numel()
        base_off = g0.storage_offset() - self.layout[0][0]
        storage = g0.untyped_storage()
        if base_off < 0 or storage.nbytes() < 4 * (base_off + total):
            return None
        base_ptr = storage.data_ptr() + 4 * base_off
        for p, (off, _, _) in zip(self.params, self.layout):
            g = p.grad
            if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != base_ptr + 4 * off:
                return None
        return g0.as_strided((total,), (1,), base_off)
