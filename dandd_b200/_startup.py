"""Start-up trims for the command-line entry points (a `dandd` run at config-2 scale is ~0.3 s of
work inside ~5 s of interpreter, torch and CUDA start-up)."""
import os

TRIM_REQUESTED = False   # set by the command-line launchers; honoured by dandd_b200.store.get_store()


def request_trim() -> None:
    """Ask for trim_torch_cuda_init() right before the store starts CUDA -- without importing torch now."""
    global TRIM_REQUESTED
    TRIM_REQUESTED = True


def trim_torch_cuda_init() -> bool:
    """torch queues, for its lazy CUDA initialisation, the registration of two sparse-BSR Triton
    operators; doing so imports triton and costs about a second.  Nothing on this path uses them,
    so the command-line tools drop that queued call.  Library users are not affected (only the
    launchers call this); DANDD_B200_KEEP_TORCH_INIT=1 keeps torch's behaviour; any surprise in
    torch's internals turns this into a no-op."""
    if os.environ.get("DANDD_B200_KEEP_TORCH_INIT") == "1":
        return False
    try:
        import torch.cuda as tc
        fn = getattr(tc, "_register_triton_kernels", None)
        queued = getattr(tc, "_queued_calls", None)
        if fn is None or not isinstance(queued, list) or tc.is_initialized():
            return False
        kept = [c for c in queued if not (isinstance(c, tuple) and c and c[0] is fn)]
        dropped = len(kept) != len(queued)
        queued[:] = kept
        return dropped
    except Exception:
        return False
