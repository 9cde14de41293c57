"""ctypes binding of libflowtimes.so (the C ABI in include/flowtimes.h).

PyTorch is used for device memory and streams only: every wrapper takes torch
CUDA tensors, checks layout/dtype, and passes raw device pointers plus the
current stream to the library.  There is NO fallback: if the shared library is
missing or a tensor is not a contiguous CUDA tensor the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import torch

FTN_F32, FTN_BF16 = 0, 1
FTN_ACT_GELU, FTN_ACT_RELU = 0, 1
ABI_VERSION = 17
FTN_MAX_K = 16
FTN_MAX_BRANCH = 8

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("FLOWTIMES_LIB", _PKG.parent / "lib" / "libflowtimes.so"))


class FtnPeriodPlan(C.Structure):
    _fields_ = [
        ("seq_len", C.c_int32),
        ("n_raw", C.c_int32),
        ("n_valid", C.c_int32),
        ("n_groups", C.c_int32),
        ("total_rows_per_window", C.c_int32),
        ("reserved", C.c_int32 * 3),
        ("raw_freq", C.c_int64 * FTN_MAX_K),
        ("freq", C.c_int64 * FTN_MAX_K),
        ("period", C.c_int64 * FTN_MAX_K),
        ("mapping", C.c_int32 * FTN_MAX_K),
        ("grp_period", C.c_int32 * FTN_MAX_K),
        ("grp_pad", C.c_int32 * FTN_MAX_K),
        ("grp_cycles", C.c_int32 * FTN_MAX_K),
        ("grp_canon", C.c_int32 * FTN_MAX_K),
        ("grp_row_off", C.c_int32 * (FTN_MAX_K + 1)),
    ]


class FtnInceptionWeights(C.Structure):
    _fields_ = [
        ("cin", C.c_int32), ("cout", C.c_int32), ("mid", C.c_int32), ("n_branch", C.c_int32),
        ("kh", C.c_int32 * FTN_MAX_BRANCH), ("kw", C.c_int32 * FTN_MAX_BRANCH),
        ("kk_cin", C.c_int32), ("kk_cout", C.c_int32),
        ("w_in", C.c_void_p), ("b_in", C.c_void_p),
        ("w_kk", C.c_void_p * FTN_MAX_BRANCH), ("b_kk", C.c_void_p * FTN_MAX_BRANCH),
        ("w_out", C.c_void_p), ("b_out", C.c_void_p),
        ("w_res", C.c_void_p), ("b_res", C.c_void_p),
        ("w_in_bf16", C.c_void_p), ("w_out_bf16", C.c_void_p), ("w_res_bf16", C.c_void_p),
        ("w_kk_bf16", C.c_void_p * FTN_MAX_BRANCH),
        ("w_mid_first", C.c_void_p), ("w_mid_second", C.c_void_p),
        ("w_kk_phase", C.c_void_p * FTN_MAX_BRANCH),
        ("w_kk_img", C.c_void_p * FTN_MAX_BRANCH), ("w_kk_img3", C.c_void_p * FTN_MAX_BRANCH),
        ("w_in_s3", C.c_void_p), ("w_out_s3", C.c_void_p), ("w_res_s3", C.c_void_p),
        ("w_in_h2", C.c_void_p), ("w_out_h2", C.c_void_p), ("w_res_h2", C.c_void_p),
        ("w_kk_img2", C.c_void_p * FTN_MAX_BRANCH),
        ("sc_in", C.c_float), ("sc_out", C.c_float), ("sc_res", C.c_float),
        ("sc_kk", C.c_float * FTN_MAX_BRANCH),
        ("w_kk_row", C.c_void_p * FTN_MAX_BRANCH), ("w_kk_row2", C.c_void_p * FTN_MAX_BRANCH),
    ]


PLAN_BYTES = C.sizeof(FtnPeriodPlan)

# name -> (restype, argtypes); must list every symbol include/flowtimes.h declares
_P, _I, _F, _SZ, _I64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64
SIGNATURES = {
    "ftn_version": (_I, []),
    "ftn_last_error": (C.c_char_p, []),
    "ftn_device_info": (_I, [C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "ftn_launch_count": (C.c_longlong, []),
    "ftn_timing_enable": (_I, [_I]),
    "ftn_timing_read": (_I, [_I, C.POINTER(C.c_double), C.POINTER(_I)]),
    "ftn_spectrum_workspace_bytes": (_SZ, [_I, _I, _I]),
    "ftn_spectrum": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _SZ, _P]),
    "ftn_select_periods": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "ftn_period_search": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _SZ, _P, _P, _P]),
    "ftn_dft_basis_bytes": (_SZ, [_I]),
    "ftn_debug_dft_trace": (_I, [C.POINTER(C.c_ulonglong)]),
    "ftn_dft_basis_build": (_I, [_I, _P, _SZ, _P]),
    "ftn_peer_create": (_I, [_I, _I, C.POINTER(_P), C.c_char_p]),
    "ftn_peer_connect": (_I, [_P, C.c_char_p]),
    "ftn_peer_allreduce": (_I, [_P, _P, _I, _P]),
    "ftn_peer_destroy": (_I, [_P]),
    "ftn_plan_build_host": (_I, [C.POINTER(_I64), _I, _I, _I, _I, C.POINTER(FtnPeriodPlan)]),
    "ftn_group_weights": (_I, [_P, _I, _I, _I, _I, _P, _P, _P]),
    "ftn_inception_workspace_bytes": (_SZ, [_I, _I, _I, C.POINTER(FtnInceptionWeights), C.POINTER(FtnInceptionWeights)]),
    "ftn_debug_tc_linear": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "ftn_debug_tc_linear_split": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "ftn_debug_tc_linear_h2": (_I, [_P, _P, _F, _P, _I, _I, _I, _P, _P, _P]),
    "ftn_debug_conv_tiled": (_I, [_P, _P, _I, _P, _I, _I, _I, C.POINTER(FtnInceptionWeights), _I, _P]),
    "ftn_period_conv": (_I, [_P, _I, _I, _I, _I, _P, _I, C.POINTER(FtnInceptionWeights),
                             C.POINTER(FtnInceptionWeights), _I, _P, _P, _SZ, _P]),
    "ftn_timesblock_fused": (_I, [_P, _I, _I, _I, _I, _P, _I, C.POINTER(FtnInceptionWeights),
                                  C.POINTER(FtnInceptionWeights), _I, _P, _P, _P, _F, _P, _P, _SZ, _P]),
    "ftn_timesblock_forward": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _SZ, _P,
                                    C.POINTER(FtnInceptionWeights), C.POINTER(FtnInceptionWeights), _I, _P, _P, _F, _P, _P,
                                    _SZ, _P, _P]),
    "ftn_inception_block_workspace_bytes": (_SZ, [_I, _I, _I, C.POINTER(FtnInceptionWeights)]),
    "ftn_inception_block": (_I, [_P, _I, _I, _I, _P, _I, C.POINTER(FtnInceptionWeights), _I, _I, _P, _P, _SZ, _P]),
    "ftn_conv2d_grid": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P]),
    "ftn_rms_norm": (_I, [_P, _I, _I, _I, _P, _P, _F, _P, _P]),
    "ftn_recursive_advance": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P]),
    "ftn_aggregate": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _F, _P, _P]),
    "ftn_context_add": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "ftn_linear": (_I, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "ftn_layer_norm": (_I, [_P, _I, _I, _I, _P, _P, _F, _P, _P]),
    "ftn_embed_combine": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "ftn_nb_head": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ftn_nb_nll_backward": (_I, [_P, _P, _P, _P, _I64, _F, _P, _P, _P, _P, _P]),
    "ftn_nb_head_epilogue_backward": (_I, [_P, _P, _P, _P, _P, _I64, _I, _P, _P, _P]),
    "ftn_layer_norm_backward": (_I, [_P, _P, _P, _I64, _I, _F, _P, _P, _P, _P]),
    "ftn_gemm_f32": (_I, [_P, _I, _I64, _I, _P, _I, _I64, _I, _P, _I, _I64, _I, _I, _I, _I, _I, _P]),
    "ftn_act_forward": (_I, [_P, _I64, _I, _P, _P]),
    "ftn_act_backward": (_I, [_P, _P, _I64, _I, _P, _P]),
    "ftn_conv2d_grid_backward_weight": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "ftn_aggregate_backward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "ftn_group_weights_backward": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "ftn_spectrum_amp_backward": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    "ftn_embed_tc_workspace_bytes": (_SZ, [C.c_longlong, _I]),
    "ftn_embed_tc": (_I, [_P, C.c_longlong, _I, _I, _P, _P, _P, _I, _P, _I, _I, _P, _P, _SZ, _P]),
    "ftn_nb_head_tc_workspace_bytes": (_SZ, [_I, _I, _I]),
    "ftn_nb_head_tc": (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "ftn_time_proj_pack_bytes": (_SZ, [_I, _I]),
    "ftn_time_proj_pack": (_I, [_P, _I, _I, _P, _SZ, _P]),
    "ftn_nb_nll": (_I, [_P, _P, _P, _P, _I64, _F, _P, _P, _P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"libflowtimes.so not found at {LIB_PATH}. Build it with "
            "`python flow-timesnet_b200/build.py` (nvcc, sm_100a). There is no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI drift, fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.ftn_version() != ABI_VERSION:
        raise RuntimeError(f"libflowtimes ABI version {lib.ftn_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


class FlowTimesError(RuntimeError):
    pass


_SYNC_CHECK = bool(os.environ.get("FLOWTIMES_SYNC_CHECK"))   # diagnostic: synchronise after every call and name the one that faulted


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ftn_last_error().decode("utf-8", "replace")
        raise FlowTimesError(f"{what} failed (rc={rc}): {msg}")
    if _SYNC_CHECK and not torch.cuda.is_current_stream_capturing():
        try:
            torch.cuda.synchronize()
        except RuntimeError as e:
            raise FlowTimesError(f"{what}: device fault after the call: {e}") from e


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return FTN_F32
    if dt == torch.bfloat16:
        return FTN_BF16
    raise TypeError(f"flowtimes supports float32 and bfloat16 activations, got {dt} (no fp16 path)")


def require_cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: the B200-native TimesBlock path runs on CUDA only (no CPU fallback)")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def device_info():
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    _check(load().ftn_device_info(C.byref(sm), C.byref(ma), C.byref(mi)), "ftn_device_info")
    return sm.value, ma.value, mi.value


def launch_count() -> int:
    return int(load().ftn_launch_count())


FAM_SPECTRUM, FAM_CONV, FAM_AGGREGATE = 0, 1, 2
FAM_S1, FAM_KK_A, FAM_MID, FAM_KK_B, FAM_S6 = 3, 4, 5, 6, 7     # single kernels of the bf16 chain
FAM_FFT, FAM_MEDIAN, FAM_SELECT = 8, 9, 10                     # single kernels of the period search


def timing_enable(on: bool) -> None:
    _check(load().ftn_timing_enable(int(on)), "ftn_timing_enable")


def timing_read(family: int):
    ms, n = C.c_double(), C.c_int()
    _check(load().ftn_timing_read(family, C.byref(ms), C.byref(n)), "ftn_timing_read")
    return ms.value, n.value


# --------------------------------------------------------------------------- #
# K1
# --------------------------------------------------------------------------- #
def spectrum(x: torch.Tensor):
    """x[B,L,C] -> (amp_median[B,F] fp32, amp_sum[F+1] fp32; the last slot is the window count B)."""
    lib = load()
    B, L, Cc = x.shape
    Fq = L // 2 + 1
    med = torch.empty(B, Fq, dtype=torch.float32, device=x.device)
    ssum = torch.empty(Fq + 1, dtype=torch.float32, device=x.device)
    nbytes = lib.ftn_spectrum_workspace_bytes(B, L, Cc)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _check(lib.ftn_spectrum(x.data_ptr(), dtype_code(x.dtype), B, L, Cc, med.data_ptr(), ssum.data_ptr(),
                            ws.data_ptr(), nbytes, _stream()), "ftn_spectrum")
    return med, ssum


PEER_HANDLE_BYTES = 64


class PeerComm:
    """NVLink peer mailbox of one rank (csrc/peer.cuh): the path's one collective -- the sum of L/2 + 2 floats over the
    ranks -- done inside the selection kernel with plain stores / loads to peer memory instead of an NCCL all-reduce.
    ``PeerComm(rank, world, gather)``: ``gather(bytes) -> list of every rank's bytes`` exchanges the CUDA IPC handles
    (torch.distributed, any backend)."""

    def __init__(self, rank: int, world: int, gather):
        lib = load()
        self.rank, self.world = int(rank), int(world)
        self._handle = C.c_void_p()
        buf = C.create_string_buffer(PEER_HANDLE_BYTES)
        _check(lib.ftn_peer_create(self.rank, self.world, C.byref(self._handle), buf), "ftn_peer_create")
        handles = gather(bytes(buf.raw))
        if len(handles) != self.world or any(len(h) != PEER_HANDLE_BYTES for h in handles):
            raise RuntimeError("PeerComm: the handle exchange did not return one 64-byte handle per rank")
        _check(lib.ftn_peer_connect(self._handle, b"".join(handles)), "ftn_peer_connect")

    @property
    def ptr(self) -> int:
        return self._handle.value

    def all_reduce(self, vals: torch.Tensor) -> torch.Tensor:
        """In place: vals (fp32, <= 1024 elements) <- sum over ranks in rank order.  Every rank must call."""
        if vals.dtype != torch.float32 or not vals.is_cuda or not vals.is_contiguous():
            raise TypeError("PeerComm.all_reduce takes a contiguous fp32 CUDA tensor")
        _check(load().ftn_peer_allreduce(self._handle, vals.data_ptr(), vals.numel(), _stream()), "ftn_peer_allreduce")
        return vals

    def close(self) -> None:
        if self._handle:
            load().ftn_peer_destroy(self._handle)
            self._handle = C.c_void_p()


_DFT_BASIS = {}   # (device index, L) -> uint8 tensor holding the three-plane DFT basis of the tensor-core spectrum


def dft_basis(x: torch.Tensor) -> Optional[torch.Tensor]:
    """DFT basis for the tensor-core spectrum of ``x[B, L, C]`` (csrc/tc_dft.cu), built once per (device, L) and cached;
    ``None`` when that route does not apply (fp32 activations, C other than 64 / 128, short windows) or while a CUDA
    graph is being captured before the basis exists (the SIMT FFT is captured instead)."""
    B, L, Cc = x.shape
    if x.dtype != torch.bfloat16 or Cc not in (64, 128) or L <= 64 or os.environ.get("FLOWTIMES_NO_TC_DFT"):
        return None
    key = (x.device.index, L)
    t = _DFT_BASIS.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            return None
        lib = load()
        nbytes = lib.ftn_dft_basis_bytes(L)
        t = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _check(lib.ftn_dft_basis_build(L, t.data_ptr(), nbytes, _stream()), "ftn_dft_basis_build")
        torch.cuda.current_stream(x.device).synchronize()   # one-time: later calls may come from any stream
        _DFT_BASIS[key] = t
    return t


def period_search(x: torch.Tensor, k: int, pmax: int, min_period: int, comm: Optional["PeerComm"] = None,
                  tensor_dft: bool = True):
    """Fused search: x[B,L,C] -> (plan, amps[B,k], weights[B,16], amp_median[B,F], amp_sum[F+1]).  With ``comm`` the
    batch is sharded over the communicator's ranks and the partial sums are exchanged inside the selection kernel.
    ``tensor_dft=False`` keeps the spectrum on the SIMT FFT (A/B tests)."""
    lib = load()
    basis = dft_basis(x) if tensor_dft else None
    B, L, Cc = x.shape
    Fq = L // 2 + 1
    med = torch.empty(B, Fq, dtype=torch.float32, device=x.device)
    ssum = torch.empty(Fq + 1, dtype=torch.float32, device=x.device)
    plan = new_plan(x.device)
    amps = torch.empty(B, k, dtype=x.dtype, device=x.device)
    weights = torch.empty(B, FTN_MAX_K, dtype=torch.float32, device=x.device)
    nbytes = lib.ftn_spectrum_workspace_bytes(B, L, Cc)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _check(lib.ftn_period_search(x.data_ptr(), dtype_code(x.dtype), B, L, Cc, k, pmax, min_period, med.data_ptr(),
                                 ssum.data_ptr(), plan.data_ptr(), amps.data_ptr(), weights.data_ptr(), ws.data_ptr(),
                                 nbytes, _ptr(basis), None if comm is None else comm.ptr, _stream()), "ftn_period_search")
    return plan, amps, weights, med, ssum


def new_plan(device) -> torch.Tensor:
    return torch.zeros(PLAN_BYTES, dtype=torch.uint8, device=device)


def select_periods(med: torch.Tensor, ssum: torch.Tensor, dtype: torch.dtype, global_batch: int, L: int, k: int,
                   pmax: int, min_period: int):
    """-> (plan uint8[PLAN_BYTES] device, amps[B,k] dtype, weights[B,FTN_MAX_K] fp32)."""
    lib = load()
    B = med.shape[0]
    plan = new_plan(med.device)
    amps = torch.empty(B, k, dtype=dtype, device=med.device)
    weights = torch.empty(B, FTN_MAX_K, dtype=torch.float32, device=med.device)
    _check(lib.ftn_select_periods(med.data_ptr(), ssum.data_ptr(), dtype_code(dtype), B, int(global_batch), L, k,
                                  pmax, min_period, plan.data_ptr(), amps.data_ptr(), weights.data_ptr(),
                                  _stream()), "ftn_select_periods")
    return plan, amps, weights


def plan_build_host(periods, L: int, min_period: Optional[int], max_period: Optional[int]) -> FtnPeriodPlan:
    lib = load()
    k = len(periods)
    arr = (C.c_int64 * max(k, 1))(*[int(p) for p in periods])
    plan = FtnPeriodPlan()
    _check(lib.ftn_plan_build_host(arr, k, int(L), int(min_period or 0), int(max_period or 0), C.byref(plan)),
           "ftn_plan_build_host")
    return plan


def plan_to_device(plan: FtnPeriodPlan, device) -> torch.Tensor:
    raw = bytes(memoryview(plan))
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device)


def plan_to_host(plan_dev: torch.Tensor) -> FtnPeriodPlan:
    raw = plan_dev.cpu().numpy().tobytes()          # one small D2H + sync
    return FtnPeriodPlan.from_buffer_copy(raw)


def group_weights(amps: torch.Tensor, plan_dev: torch.Tensor, B: int) -> torch.Tensor:
    lib = load()
    if amps.dim() == 1:
        amps = amps.view(1, -1)
    k = amps.shape[1]
    stride = k if amps.shape[0] == B else 0
    if amps.shape[0] not in (1, B):
        raise ValueError("amplitudes must have shape [B, K] or [K]")
    weights = torch.empty(B, FTN_MAX_K, dtype=torch.float32, device=amps.device)
    _check(lib.ftn_group_weights(amps.data_ptr(), dtype_code(amps.dtype), B, k, stride, plan_dev.data_ptr(),
                                 weights.data_ptr(), _stream()), "ftn_group_weights")
    return weights


# --------------------------------------------------------------------------- #
# K2-K4
# --------------------------------------------------------------------------- #
def inception_workspace_bytes(B: int, L: int, max_groups: int, wa: FtnInceptionWeights, wb: FtnInceptionWeights) -> int:
    return int(load().ftn_inception_workspace_bytes(B, L, max_groups, C.byref(wa), C.byref(wb)))


def period_conv(x: torch.Tensor, plan_dev: torch.Tensor, max_groups: int, wa: FtnInceptionWeights,
                wb: FtnInceptionWeights, act: int, delta: torch.Tensor, ws: torch.Tensor) -> None:
    B, L, Cc = x.shape
    _check(load().ftn_period_conv(x.data_ptr(), dtype_code(x.dtype), B, L, Cc, plan_dev.data_ptr(), max_groups,
                                  C.byref(wa), C.byref(wb), act, delta.data_ptr(), ws.data_ptr(), ws.numel(),
                                  _stream()), "ftn_period_conv")


def timesblock_fused(x: torch.Tensor, plan_dev: torch.Tensor, max_groups: int, wa: FtnInceptionWeights,
                     wb: FtnInceptionWeights, act: int, weights: torch.Tensor, ln_w: Optional[torch.Tensor],
                     ln_b: Optional[torch.Tensor], eps: float, out: torch.Tensor, ws: torch.Tensor) -> bool:
    """Fused K2+K3+K4.  Returns False when the configuration is not eligible (caller runs the unfused pair)."""
    B, L, Cc = x.shape
    rc = load().ftn_timesblock_fused(x.data_ptr(), dtype_code(x.dtype), B, L, Cc, plan_dev.data_ptr(), max_groups,
                                     C.byref(wa), C.byref(wb), act, weights.data_ptr(), _ptr(ln_w), _ptr(ln_b),
                                     float(eps), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    if rc == -1:
        return False
    _check(rc, "ftn_timesblock_fused")
    return True


def timesblock_forward(x: torch.Tensor, k: int, pmax: int, min_period: int, wa: FtnInceptionWeights,
                       wb: FtnInceptionWeights, act: int, ln_w: Optional[torch.Tensor], ln_b: Optional[torch.Tensor],
                       eps: float, comm: Optional["PeerComm"] = None, plan: Optional[torch.Tensor] = None):
    """Period search + fused TimesBlock in one call (the first 1x1 stage overlaps the search).

    ``plan``: a plan buffer to reuse (zero before its FIRST use -- every search leaves its ticket zero and rewrites the
    rest), which saves the fill kernel of a fresh ``new_plan`` per call.
    Returns None when the configuration is not eligible (nothing was enqueued), else
    (out, plan, amps[B,k], weights[B,16])."""
    lib = load()
    B, L, Cc = x.shape
    Fq = L // 2 + 1
    med = torch.empty(B, Fq, dtype=torch.float32, device=x.device)
    ssum = torch.empty(Fq + 1, dtype=torch.float32, device=x.device)
    if plan is None:
        plan = new_plan(x.device)
    amps = torch.empty(B, k, dtype=x.dtype, device=x.device)
    weights = torch.empty(B, FTN_MAX_K, dtype=torch.float32, device=x.device)
    sbytes = lib.ftn_spectrum_workspace_bytes(B, L, Cc)
    sws = torch.empty(sbytes, dtype=torch.uint8, device=x.device)
    basis = dft_basis(x)
    nbytes = lib.ftn_inception_workspace_bytes(B, L, k, C.byref(wa), C.byref(wb))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty_like(x)
    rc = lib.ftn_timesblock_forward(x.data_ptr(), dtype_code(x.dtype), B, L, Cc, k, pmax, min_period, med.data_ptr(),
                                    ssum.data_ptr(), plan.data_ptr(), amps.data_ptr(), weights.data_ptr(), sws.data_ptr(),
                                    sbytes, _ptr(basis), C.byref(wa), C.byref(wb), act, _ptr(ln_w), _ptr(ln_b), float(eps),
                                    out.data_ptr(), ws.data_ptr(), nbytes, None if comm is None else comm.ptr, _stream())
    if rc == -1:
        return None
    _check(rc, "ftn_timesblock_forward")
    return out, plan, amps, weights


def single_group_plan(period: int, cycles: int, device) -> torch.Tensor:
    """Device plan holding ONE group (period, cycles, no padding): the fold of an NCHW grid [B, C, cycles, period]."""
    pl = FtnPeriodPlan()
    L = int(period) * int(cycles)
    pl.seq_len, pl.n_raw, pl.n_valid, pl.n_groups, pl.total_rows_per_window = L, 1, 1, 1, L
    for i in range(FTN_MAX_K):
        pl.mapping[i] = -1
        pl.grp_canon[i] = -1
        pl.grp_row_off[i] = L
    pl.mapping[0] = 0
    pl.period[0] = int(period)
    pl.grp_period[0], pl.grp_pad[0], pl.grp_cycles[0], pl.grp_canon[0], pl.grp_row_off[0] = int(period), 0, int(cycles), 0, 0
    pl.grp_row_off[FTN_MAX_K] = L
    return plan_to_device(pl, device)


def inception_block(x: torch.Tensor, plan_dev: torch.Tensor, w: FtnInceptionWeights, act: int,
                    trailing_act: bool) -> torch.Tensor:
    """One packed InceptionBlock on x[B, L, cin] folded by a ONE-group plan -> fp32 [B, L, cout]."""
    lib = load()
    B, L, _ = x.shape
    out = torch.empty(B, L, int(w.cout), dtype=torch.float32, device=x.device)
    nbytes = lib.ftn_inception_block_workspace_bytes(B, L, 1, C.byref(w))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    _check(lib.ftn_inception_block(x.data_ptr(), dtype_code(x.dtype), B, L, plan_dev.data_ptr(), 1, C.byref(w), act,
                                   int(trailing_act), out.data_ptr(), ws.data_ptr(), nbytes, _stream()),
           "ftn_inception_block")
    return out


def conv2d_grid(x: torch.Tensor, plan_dev: torch.Tensor, w_taps: torch.Tensor, bias: torch.Tensor, kh: int,
                kw: int) -> torch.Tensor:
    """conv2d with zero "same" padding on x[B, L, cin] (fp32) folded by a ONE-group plan; w_taps [kh*kw, cin, cout]."""
    B, L, cin = x.shape
    cout = w_taps.shape[-1]
    out = torch.empty(B, L, cout, dtype=torch.float32, device=x.device)
    _check(load().ftn_conv2d_grid(x.data_ptr(), B, L, cin, cout, kh, kw, plan_dev.data_ptr(), 1, w_taps.data_ptr(),
                                  bias.data_ptr(), out.data_ptr(), _stream()), "ftn_conv2d_grid")
    return out


def debug_tc_linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    _check(load().ftn_debug_tc_linear(a.data_ptr(), w.data_ptr(), bias.data_ptr(), M, K, N, out.data_ptr(), _stream()),
           "ftn_debug_tc_linear")
    return out


def debug_tc_linear_split(a: torch.Tensor, w_s3: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """fp32 a[M, K] x split weights w_s3[N, 3K] (bf16 planes) -> fp32 [M, N] reassembled from the three output planes."""
    M, K = a.shape
    N = w_s3.shape[0]
    a_ws = torch.empty(M, 3 * K, dtype=torch.bfloat16, device=a.device)
    out = torch.empty(M, 3 * N, dtype=torch.bfloat16, device=a.device)
    _check(load().ftn_debug_tc_linear_split(a.data_ptr(), w_s3.data_ptr(), bias.data_ptr(), M, K, N, a_ws.data_ptr(),
                                            out.data_ptr(), _stream()), "ftn_debug_tc_linear_split")
    o = out.view(M, 3, N).float()
    return o[:, 0] + o[:, 1] + o[:, 2]


def debug_tc_linear_h2(a: torch.Tensor, w_h2: torch.Tensor, scale: float, bias: torch.Tensor) -> torch.Tensor:
    """fp32 a[M, K] x two-plane fp16 weights w_h2[N, 2K] (scaled by 1 / scale) -> fp32 [M, N] from the two output planes."""
    M, K = a.shape
    N = w_h2.shape[0]
    a_ws = torch.empty(M, 2 * K, dtype=torch.float16, device=a.device)
    out = torch.empty(M, 2 * N, dtype=torch.float16, device=a.device)
    _check(load().ftn_debug_tc_linear_h2(a.data_ptr(), w_h2.data_ptr(), float(scale), bias.data_ptr(), M, K, N,
                                         a_ws.data_ptr(), out.data_ptr(), _stream()), "ftn_debug_tc_linear_h2")
    o = out.view(M, 2, N).float()
    return o[:, 1] + o[:, 0]


def debug_conv_tiled(inp: torch.Tensor, plan_dev: torch.Tensor, B: int, L: int, max_groups: int,
                     w: FtnInceptionWeights, use_tc: bool) -> torch.Tensor:
    out = torch.zeros_like(inp)
    _check(load().ftn_debug_conv_tiled(inp.data_ptr(), out.data_ptr(), inp.shape[1], plan_dev.data_ptr(), B, L,
                                       max_groups, C.byref(w), int(use_tc), _stream()), "ftn_debug_conv_tiled")
    return out


def aggregate(x: torch.Tensor, delta: torch.Tensor, weights: torch.Tensor, plan_dev: torch.Tensor,
              ln_w: Optional[torch.Tensor], ln_b: Optional[torch.Tensor], eps: float, out: torch.Tensor) -> None:
    B, L, Cc = x.shape
    _check(load().ftn_aggregate(x.data_ptr(), delta.data_ptr(), weights.data_ptr(), plan_dev.data_ptr(),
                                dtype_code(x.dtype), B, L, Cc, _ptr(ln_w), _ptr(ln_b), float(eps), out.data_ptr(),
                                _stream()), "ftn_aggregate")


# --------------------------------------------------------------------------- #
# K5, K6, helpers
# --------------------------------------------------------------------------- #
def context_add(x, coeff, basis, scale, out) -> None:
    B, L, N = x.shape
    R = coeff.shape[-1]
    _check(load().ftn_context_add(x.data_ptr(), coeff.data_ptr(), basis.data_ptr(), scale.data_ptr(), B, L, N, R,
                                  out.data_ptr(), _stream()), "ftn_context_add")


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """fp32 Linear on the last dim: a[..., K] x w[N, K]^T + bias."""
    K = a.shape[-1]
    M = a.numel() // K
    N = w.shape[0]
    out = torch.empty(*a.shape[:-1], N, dtype=torch.float32, device=a.device)
    _check(load().ftn_linear(a.data_ptr(), w.data_ptr(), _ptr(bias), M, K, N, out.data_ptr(), _stream()), "ftn_linear")
    return out


def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out = torch.empty_like(x)
    _check(load().ftn_layer_norm(x.data_ptr(), dtype_code(x.dtype), rows, Cc, w.data_ptr(), b.data_ptr(), float(eps),
                                 out.data_ptr(), _stream()), "ftn_layer_norm")
    return out


def rms_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    out = torch.empty_like(x)
    _check(load().ftn_rms_norm(x.data_ptr(), dtype_code(x.dtype), rows, Cc, w.data_ptr(), b.data_ptr(), float(eps),
                               out.data_ptr(), _stream()), "ftn_rms_norm")
    return out


def recursive_advance(window, rate, disp, rates, disps, mark, y_mark, step_counter) -> None:
    B, L, N = window.shape
    H = rates.shape[1]
    Tm = 0 if mark is None else mark.shape[-1]
    _check(load().ftn_recursive_advance(window.data_ptr(), rate.data_ptr(), disp.data_ptr(), B, L, N, H, rates.data_ptr(),
                                        disps.data_ptr(), _ptr(mark), _ptr(y_mark), Tm, step_counter.data_ptr(), _stream()),
           "ftn_recursive_advance")


def embed_combine(value, aux, gate, aux_batched: bool, out_dtype: torch.dtype) -> torch.Tensor:
    B, L, Cc = value.shape
    out = torch.empty(B, L, Cc, dtype=out_dtype, device=value.device)
    _check(load().ftn_embed_combine(value.data_ptr(), aux.data_ptr(), gate.data_ptr(), int(aux_batched), B, L, Cc,
                                    dtype_code(out_dtype), out.data_ptr(), _stream()), "ftn_embed_combine")
    return out


def embed_tc(x: torch.Tensor, w_s3: torch.Tensor, bias: torch.Tensor, aux: torch.Tensor, aux_batched: bool,
             gate: torch.Tensor, out_dtype: torch.dtype) -> Optional[torch.Tensor]:
    """DataEmbedding on the tensor cores (value GEMM + combine in one kernel).  None = shape not eligible."""
    lib = load()
    B, L, N = x.shape
    Cc = w_s3.shape[0]
    rows = B * L
    out = torch.empty(B, L, Cc, dtype=out_dtype, device=x.device)
    nbytes = lib.ftn_embed_tc_workspace_bytes(rows, N)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    rc = lib.ftn_embed_tc(x.data_ptr(), rows, L, N, w_s3.data_ptr(), bias.data_ptr(), aux.data_ptr(), int(aux_batched),
                          gate.data_ptr(), Cc, dtype_code(out_dtype), out.data_ptr(), ws.data_ptr(), nbytes, _stream())
    if rc == -1:
        return None
    _check(rc, "ftn_embed_tc")
    return out


def _hist_view(hist: torch.Tensor, steps: int, N: int):
    """(pointer, batch stride in elements) of a history tail ``[B, steps, N]`` that may be a VIEW into x (rows dense)."""
    if hist.stride(2) != 1 or hist.stride(1) != N:
        hist = hist.contiguous()
    return hist, hist.data_ptr(), int(hist.stride(0)) if hist.shape[0] > 1 else steps * N


def time_proj_pack(Wt: torch.Tensor) -> torch.Tensor:
    """forecast_time_proj rows ``[steps, L]`` fp32 -> the three-plane bf16 operand of the tensor-core time projection."""
    lib = load()
    steps, L = Wt.shape
    nbytes = lib.ftn_time_proj_pack_bytes(steps, L)
    out = torch.empty(nbytes, dtype=torch.uint8, device=Wt.device)
    _check(lib.ftn_time_proj_pack(Wt.data_ptr(), steps, L, out.data_ptr(), nbytes, _stream()), "ftn_time_proj_pack")
    return out


def nb_head_tc(seq, steps, N, Wt, bt, w_heads_s3, b_heads, Np, hist, late, late_gate, floor_n, flags, wt_s3=None):
    """NB head with the mu / sigma heads as one tensor-core GEMM (and, given ``wt_s3`` from ``time_proj_pack`` and a bf16
    ``seq``, the time projection too).  None = shape not eligible."""
    lib = load()
    B, L, Cc = seq.shape
    rate = torch.empty(B, steps, N, dtype=torch.float32, device=seq.device)
    disp = torch.empty(B, steps, N, dtype=torch.float32, device=seq.device)
    nbytes = lib.ftn_nb_head_tc_workspace_bytes(B, steps, Cc)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=seq.device)
    hist, hist_ptr, hist_stride = _hist_view(hist, steps, N)
    rc = lib.ftn_nb_head_tc(seq.data_ptr(), dtype_code(seq.dtype), B, L, Cc, steps, N, Wt.data_ptr(), bt.data_ptr(),
                            _ptr(wt_s3), w_heads_s3.data_ptr(), b_heads.data_ptr(), Np, hist_ptr, hist_stride, _ptr(late), _ptr(late_gate),
                            floor_n.data_ptr(), rate.data_ptr(), disp.data_ptr(), flags.data_ptr(), ws.data_ptr(), nbytes,
                            _stream())
    if rc == -1:
        return None
    _check(rc, "ftn_nb_head_tc")
    return rate, disp


def nb_head(seq, steps, N, Wt, bt, Wmu, bmu, Wsg, bsg, hist, late, late_gate, floor_n, flags):
    B, L, Cc = seq.shape
    rate = torch.empty(B, steps, N, dtype=torch.float32, device=seq.device)
    disp = torch.empty(B, steps, N, dtype=torch.float32, device=seq.device)
    ws = torch.empty(B * steps * Cc, dtype=torch.float32, device=seq.device)
    hist, hist_ptr, hist_stride = _hist_view(hist, steps, N)
    _check(load().ftn_nb_head(seq.data_ptr(), dtype_code(seq.dtype), B, L, Cc, steps, N, Wt.data_ptr(), bt.data_ptr(),
                              Wmu.data_ptr(), bmu.data_ptr(), Wsg.data_ptr(), bsg.data_ptr(), hist_ptr, hist_stride,
                              _ptr(late), _ptr(late_gate), floor_n.data_ptr(), rate.data_ptr(), disp.data_ptr(),
                              flags.data_ptr(), ws.data_ptr(), _stream()), "ftn_nb_head")
    return rate, disp


def nb_nll(y, rate, disp, mask_u8, eps: float) -> torch.Tensor:
    out = torch.empty((), dtype=torch.float32, device=y.device)
    partial = torch.empty(2 * 1024, dtype=torch.float32, device=y.device)
    _check(load().ftn_nb_nll(y.data_ptr(), rate.data_ptr(), disp.data_ptr(), _ptr(mask_u8), y.numel(), float(eps),
                             partial.data_ptr(), out.data_ptr(), _stream()), "ftn_nb_nll")
    return out


# --------------------------------------------------------------------------- #
# backward, first slice
# --------------------------------------------------------------------------- #
def nb_nll_backward(y, rate, disp, mask_u8, eps: float, grad_out: torch.Tensor):
    d_rate = torch.empty_like(rate)
    d_disp = torch.empty_like(disp)
    scratch = torch.empty(1, dtype=torch.float32, device=y.device)
    go = grad_out.detach().to(device=y.device, dtype=torch.float32).reshape(1).contiguous()
    _check(load().ftn_nb_nll_backward(y.data_ptr(), rate.data_ptr(), disp.data_ptr(), _ptr(mask_u8), y.numel(), float(eps),
                                      go.data_ptr(), scratch.data_ptr(), d_rate.data_ptr(), d_disp.data_ptr(), _stream()),
           "ftn_nb_nll_backward")
    return d_rate, d_disp


def nb_head_epilogue_backward(rate, disp, floor_n, d_rate, d_disp):
    N = rate.shape[-1]
    rows = rate.numel() // N
    dpr = torch.empty_like(rate)
    dpd = torch.empty_like(disp)
    _check(load().ftn_nb_head_epilogue_backward(rate.data_ptr(), disp.data_ptr(), floor_n.data_ptr(), d_rate.data_ptr(),
                                                d_disp.data_ptr(), rows, N, dpr.data_ptr(), dpd.data_ptr(), _stream()),
           "ftn_nb_head_epilogue_backward")
    return dpr, dpd


def layer_norm_backward(x, dy, w, eps: float):
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    dx = torch.empty_like(x)
    dw = torch.empty(Cc, dtype=torch.float32, device=x.device)
    db = torch.empty(Cc, dtype=torch.float32, device=x.device)
    _check(load().ftn_layer_norm_backward(x.data_ptr(), dy.data_ptr(), w.data_ptr(), rows, Cc, float(eps), dx.data_ptr(),
                                          dw.data_ptr(), db.data_ptr(), _stream()), "ftn_layer_norm_backward")
    return dx, dw, db


def gemm_f32(a: torch.Tensor, b: torch.Tensor, trans_a: bool = False, trans_b: bool = False) -> torch.Tensor:
    """op(a) @ op(b) for contiguous fp32 matrices ``[.., r, c]`` with an optional shared leading batch dim."""
    batch = a.shape[0] if a.dim() == 3 else (b.shape[0] if b.dim() == 3 else 1)
    ar, ac = a.shape[-2], a.shape[-1]
    br, bc = b.shape[-2], b.shape[-1]
    M, K = (ac, ar) if trans_a else (ar, ac)
    K2, N = (bc, br) if trans_b else (br, bc)
    if K != K2:
        raise ValueError(f"gemm_f32: inner dimensions differ ({K} vs {K2})")
    shape = (batch, M, N) if (a.dim() == 3 or b.dim() == 3) else (M, N)
    out = torch.empty(shape, dtype=torch.float32, device=a.device)
    _check(load().ftn_gemm_f32(a.data_ptr(), ac, ar * ac if a.dim() == 3 else 0, int(trans_a), b.data_ptr(), bc,
                               br * bc if b.dim() == 3 else 0, int(trans_b), out.data_ptr(), N, M * N if len(shape) == 3 else 0,
                               M, N, K, batch, 0, _stream()), "ftn_gemm_f32")
    return out


# --------------------------------------------------------------------------- #
# backward, second slice (csrc/backward.cu)
# --------------------------------------------------------------------------- #
def act_forward(x: torch.Tensor, act: int) -> torch.Tensor:
    out = torch.empty_like(x)
    _check(load().ftn_act_forward(x.data_ptr(), x.numel(), act, out.data_ptr(), _stream()), "ftn_act_forward")
    return out


def act_backward(x: torch.Tensor, dy: torch.Tensor, act: int) -> torch.Tensor:
    dx = torch.empty_like(x)
    _check(load().ftn_act_backward(x.data_ptr(), dy.data_ptr(), x.numel(), act, dx.data_ptr(), _stream()), "ftn_act_backward")
    return dx


def conv2d_grid_backward_weight(x: torch.Tensor, dy: torch.Tensor, period: int, kh: int, kw: int) -> torch.Tensor:
    """dW ``[kh*kw, cin, cout]`` of ``conv2d_grid`` for x ``[B, L, cin]``, dy ``[B, L, cout]`` (fp32, one period group)."""
    B, L, cin = x.shape
    cout = dy.shape[-1]
    dw = torch.empty(kh * kw, cin, cout, dtype=torch.float32, device=x.device)
    _check(load().ftn_conv2d_grid_backward_weight(x.data_ptr(), dy.data_ptr(), B, L, int(period), cin, cout, kh, kw,
                                                  dw.data_ptr(), _stream()), "ftn_conv2d_grid_backward_weight")
    return dw


def aggregate_backward(d_out: torch.Tensor, delta: torch.Tensor, weights: torch.Tensor, plan_dev: torch.Tensor):
    """-> (d_delta ``[G_slots, B, L, C]``, d_weights ``[B, FTN_MAX_K]``); d_x of the aggregation is d_out itself."""
    B, L, Cc = d_out.shape
    d_delta = torch.zeros_like(delta)
    d_w = torch.empty(B, FTN_MAX_K, dtype=torch.float32, device=d_out.device)
    _check(load().ftn_aggregate_backward(d_out.data_ptr(), delta.data_ptr(), weights.data_ptr(), plan_dev.data_ptr(), B, L, Cc,
                                         d_delta.data_ptr(), d_w.data_ptr(), _stream()), "ftn_aggregate_backward")
    return d_delta, d_w


def group_weights_backward(amps: torch.Tensor, plan_dev: torch.Tensor, d_weights: torch.Tensor) -> torch.Tensor:
    B, k = amps.shape
    d_amps = torch.empty_like(amps)
    _check(load().ftn_group_weights_backward(amps.data_ptr(), B, k, plan_dev.data_ptr(), d_weights.data_ptr(),
                                             d_amps.data_ptr(), _stream()), "ftn_group_weights_backward")
    return d_amps


def spectrum_amp_backward(x: torch.Tensor, plan_dev: torch.Tensor, d_amps: torch.Tensor) -> torch.Tensor:
    """d_x ``[B, L, C]`` of the per-window amplitudes at the plan's bins (gradient to the lower-median channel)."""
    B, L, Cc = x.shape
    d_x = torch.zeros_like(x)
    _check(load().ftn_spectrum_amp_backward(x.data_ptr(), B, L, Cc, d_amps.shape[1], plan_dev.data_ptr(), d_amps.data_ptr(),
                                            d_x.data_ptr(), _stream()), "ftn_spectrum_amp_backward")
    return d_x
