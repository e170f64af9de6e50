"""Dry-run harness for the CPU test tier: runs the HOST side of the CUDA path (ops wrappers, launch schedules, autograd
glue, module classes) on CPU tensors with every kernel-launching entry point of libw2v2_b200.so replaced by a stub
that only validates its argument list against the ctypes signature table.  Nothing is computed -- output buffers keep
whatever ``torch.empty`` gave them -- so this checks control flow, shapes, buffer plumbing and the C call sequence,
never numbers (the `-m gpu` tests do that).  The pure host-side entry points (workspace sizes, tile heuristics) are
served by the real library, which loads without a GPU."""
import contextlib
import ctypes

import torch

# entry points that run on the host only (no kernel launch, no CUDA context)
HOST_ONLY = {"w2v2_last_error", "w2v2_abi_version", "w2v2_conv0_workspace_bytes", "w2v2_conv0_workspace_offsets",
             "w2v2_posconv_taps_per_mma", "w2v2_launch_count", "w2v2_reset_launch_count", "w2v2_dgrad_accumulates",
             "w2v2_prepare_tile_edge"}


class DryLib:
    def __init__(self, real, signatures):
        self._real, self._sig = real, signatures
        self.calls = []

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        if name in HOST_ONLY:
            return getattr(self._real, name)
        res, argtypes = self._sig[name]                       # KeyError: an entry point the binding table lacks

        def stub(*args):
            assert len(args) == len(argtypes), f"{name}: {len(args)} arguments for {len(argtypes)} parameters"
            for i, (a, t) in enumerate(zip(args, argtypes)):
                try:
                    t.from_param(a)
                except (ctypes.ArgumentError, TypeError) as e:
                    raise AssertionError(f"{name}: argument {i} = {a!r} does not convert to {t.__name__}") from e
            self.calls.append(name)
            return 0

        return stub


class _DryStream:
    """torch.cuda.Stream stand-in for the dry run: ordering calls are recorded, nothing waits."""
    log = []

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        _DryStream.log.append(("wait", id(self), id(other)))

    def synchronize(self):
        pass


@contextlib.contextmanager
def dry_library():
    """Patch the package so that its CUDA path 'runs' on CPU tensors; yields the DryLib (``.calls`` = the sequence of
    launching entry points that were invoked)."""
    from w2v2_speaker_b200 import _lib, ops, schedule
    real = _lib.load()
    dry = DryLib(real, _lib.SIGNATURES)
    saved = (_lib.load, _lib.stream_ptr, ops.stream_ptr, schedule._lib.stream_ptr, torch.Tensor.is_cuda)
    cuda_saved = (torch.cuda.Stream, torch.cuda.current_stream, torch.cuda.stream)
    _lib.load = lambda: dry
    _lib.stream_ptr = ops.stream_ptr = lambda: None
    torch.Tensor.is_cuda = property(lambda self: True)          # the wrappers' "CUDA tensors only" guards
    current = _DryStream()
    torch.cuda.Stream = _DryStream                              # trainer.py: optimizer / communication streams
    torch.cuda.current_stream = lambda *a, **k: current
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    try:
        yield dry
    finally:
        _lib.load, _lib.stream_ptr, ops.stream_ptr = saved[0], saved[1], saved[2]
        torch.Tensor.is_cuda = saved[4]
        torch.cuda.Stream, torch.cuda.current_stream, torch.cuda.stream = cuda_saved
