"""An NCCL communicator for the C library's data-parallel step ("NCCL communicator passed in as an opaque pointer").

`ntf_fnn_step` sums the gradient arena over the ranks itself (two overlapped all-reduces on its own stream, capturable in a CUDA graph),
so it needs a raw `ncclComm_t` and the address of `ncclAllReduce`.  torch.distributed keeps its communicators private; this module
creates one more on the same ranks with the NCCL that torch already loaded (ctypes, no build-time dependency): rank 0 draws the unique
id, the existing process group broadcasts its 128 bytes, every rank calls `ncclCommInitRank`.  torch.distributed stays the plumbing
(rendezvous, barriers, loss / metric reductions); nothing here computes.
"""
import ctypes as C
import glob
import os

import torch


class _UniqueId(C.Structure):
    _fields_ = [('internal', C.c_byte * 128)]  # ncclUniqueId (NCCL_UNIQUE_ID_BYTES)


_nccl = None


def _load():
    global _nccl
    if _nccl is not None: return _nccl
    names = ['libnccl.so.2']  # the soname torch's bundled NCCL is already loaded under
    for base in {os.path.dirname(os.path.dirname(torch.__file__))}: names += sorted(glob.glob(os.path.join(base, 'nvidia', 'nccl', 'lib', 'libnccl.so*')))
    err = None
    for n in names:
        try:
            lib = C.CDLL(n, mode=C.RTLD_GLOBAL)
            break
        except OSError as e: err = e
    else: raise RuntimeError(f'libnccl not found ({err})')
    lib.ncclGetUniqueId.argtypes, lib.ncclGetUniqueId.restype = [C.POINTER(_UniqueId)], C.c_int
    lib.ncclCommInitRank.argtypes, lib.ncclCommInitRank.restype = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int], C.c_int
    lib.ncclCommDestroy.argtypes, lib.ncclCommDestroy.restype = [C.c_void_p], C.c_int
    lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.ncclAllReduce.restype = C.c_int
    lib.ncclGetErrorString.argtypes, lib.ncclGetErrorString.restype = [C.c_int], C.c_char_p
    _nccl = lib
    return lib


def _check(rc, what):
    if rc != 0: raise RuntimeError(f'{what} failed: {_load().ncclGetErrorString(rc).decode()} (ncclResult {rc})')


class Comm:
    """comm.ptr = ncclComm_t, comm.allreduce_addr = &ncclAllReduce (what ntf_fnn_step_args.comm / .allreduce take)"""

    def __init__(self, rank, world, device):
        lib = _load()
        self.rank, self.world, self.device = int(rank), int(world), torch.device(device)
        uid = _UniqueId()
        if self.rank == 0: _check(lib.ncclGetUniqueId(C.byref(uid)), 'ncclGetUniqueId')
        if self.world > 1:  # the id travels through the process group that already exists (one process per GPU, torchrun)
            t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=self.device if torch.distributed.get_backend() == 'nccl' else 'cpu')
            torch.distributed.broadcast(t, src=0)
            C.memmove(C.byref(uid), bytes(t.cpu().tolist()), 128)
        comm = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(lib.ncclCommInitRank(C.byref(comm), self.world, uid, self.rank), 'ncclCommInitRank')
        self.ptr = comm.value
        self.allreduce_addr = C.cast(lib.ncclAllReduce, C.c_void_p).value
        # connect now: the first collective sets up transports and buffers, which must not happen inside a stream capture
        warm = torch.zeros(1 << 20, dtype=torch.float32, device=self.device)
        self.allreduce(warm)
        torch.cuda.synchronize(self.device)

    def allreduce(self, t):
        """in-place fp32 SUM over the ranks on torch's current stream (the eager counterpart of what the step does itself)"""
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float32
        st = torch.cuda.current_stream(self.device).cuda_stream
        _check(_load().ncclAllReduce(t.data_ptr(), t.data_ptr(), t.numel(), 7, 0, self.ptr, st), 'ncclAllReduce')
        return t

    def destroy(self):
        if getattr(self, 'ptr', None):
            try: _load().ncclCommDestroy(self.ptr)
            except Exception: pass
            self.ptr = None
