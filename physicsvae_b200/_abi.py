"""ctypes binding of libpvae_sm100.so (the C ABI of include/pvae_sm100.h).

This is the only place the package touches native code.  There is no CPU fallback: if the shared library is missing
or no B200 is present every compute entry point raises.  The oracle under oracle/ is test infrastructure and is never
imported from here.
"""
import ctypes as C
import os

PVAE_ABI_VERSION = 2
PVAE_MAX_LAYERS = 8
PVAE_NUM_NETS = 4
PVAE_LOSS_SLOTS = 8

NET_TASK_ENCODER, NET_MOTOR_DECODER, NET_WORLD_MODEL, NET_VALUE_BRANCH = 0, 1, 2, 3
PREC_BF16, PREC_BF16X3 = 1, 3
PART_ENCODER, PART_DECODER, PART_WORLD, PART_VALUE = 1, 2, 4, 8

# activation registry of the reference: get_activation_fn, rllib_model_torch.py:30-46
ACT_IDS = {None: 0, "linear": 0, "relu": 1, "tanh": 2, "sigmoid": 3, "elu": 4, "swish": 5, "silu": 5}

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
# PVAE_LIB selects another build of the same library (e.g. one compiled with -DPVAE_DEBUG_HOOKS for the role-timeline tools)
LIB_PATH = os.environ.get("PVAE_LIB") or os.path.join(LIB_DIR, "libpvae_sm100.so")

# every symbol include/pvae_sm100.h declares (tests/test_abi.py checks the library exports exactly these)
SYMBOLS = [
    "pvae_last_error", "pvae_abi_version", "pvae_create", "pvae_destroy", "pvae_bind_net", "pvae_net_grad_elems",
    "pvae_sync_weights", "pvae_adam_step", "pvae_workspace_bytes", "pvae_bind_workspace", "pvae_transitions_bytes", "pvae_ingest", "pvae_ingest_episodes",
    "pvae_bind_transitions", "pvae_set_cursor", "pvae_advance_cursor", "pvae_world_step", "pvae_vae_step",
    "pvae_forward", "pvae_gemm_bf16", "pvae_launch_count", "pvae_debug_trace", "pvae_eval_loss", "pvae_noise_counter", "pvae_fc_forward",
    "pvae_symm_allreduce", "pvae_symm_flag_elems", "pvae_rollout_workspace_bytes", "pvae_bind_rollout_workspace", "pvae_rollout_step", "pvae_set_deterministic", "pvae_set_exchange", "pvae_run_exchange", "pvae_bind_act_params",
]


class NetDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("out_dims", C.c_int32 * PVAE_MAX_LAYERS), ("acts", C.c_int32 * PVAE_MAX_LAYERS),
                ("in_dims", C.c_int32 * 2)]


class ModelDesc(C.Structure):
    _fields_ = [("dim_state_body", C.c_int32), ("dim_action", C.c_int32), ("latent_dim", C.c_int32),
                ("latent_prior", C.c_int32), ("precision", C.c_int32), ("max_batch", C.c_int32),
                ("nets", NetDesc * PVAE_NUM_NETS)]


class PvaeError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once) and declare prototypes.  Raises PvaeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PvaeError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_size_t
    lib.pvae_last_error.restype = C.c_char_p
    lib.pvae_last_error.argtypes = []
    lib.pvae_abi_version.restype = i32
    lib.pvae_abi_version.argtypes = []
    lib.pvae_create.argtypes = [C.POINTER(vp), C.POINTER(ModelDesc), i32]
    lib.pvae_destroy.argtypes = [vp]
    lib.pvae_bind_net.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), vp]
    lib.pvae_bind_act_params.argtypes = [vp, i32, vp, vp]
    lib.pvae_net_grad_elems.restype = i64
    lib.pvae_net_grad_elems.argtypes = [vp, i32]
    lib.pvae_sync_weights.argtypes = [vp, u32, vp]
    lib.pvae_adam_step.argtypes = [vp, i32, u32, vp, vp, vp, f32, f32, f32, f32, f32, vp]
    lib.pvae_workspace_bytes.argtypes = [vp, C.POINTER(sz)]
    lib.pvae_bind_workspace.argtypes = [vp, vp, sz]
    lib.pvae_transitions_bytes.argtypes = [vp, i64, C.POINTER(sz)]
    lib.pvae_ingest.argtypes = [vp, vp, i64, i64, vp, i32, vp, i64, vp]
    lib.pvae_ingest_episodes.argtypes = [vp, vp, i64, i64, vp, i32, i64, vp, vp, i64, vp]
    lib.pvae_bind_transitions.argtypes = [vp, vp, i64]
    lib.pvae_set_cursor.argtypes = [vp, i64, vp]
    lib.pvae_advance_cursor.argtypes = [vp, i64, i64, i64, vp]
    lib.pvae_world_step.argtypes = [vp, i32, f32, vp, vp]
    lib.pvae_vae_step.argtypes = [vp, i32, vp, u64, u64, i32, f32, f32, f32, vp, vp]
    lib.pvae_forward.argtypes = [vp, u32, i32, vp, i64, vp, vp, i64, vp, i32, u64, u64, vp, i64, vp, vp, vp, vp, vp, vp]
    lib.pvae_eval_loss.argtypes = [vp, i32, i32, vp, u64, u64, i32, f32, f32, f32, f32, vp, vp]
    lib.pvae_noise_counter.argtypes = [vp, i32, u64, u64, vp]
    lib.pvae_fc_forward.argtypes = [vp, i32, i32, vp, i64, vp, i64, vp]
    lib.pvae_rollout_workspace_bytes.argtypes = [vp, i32, C.POINTER(sz)]
    lib.pvae_bind_rollout_workspace.argtypes = [vp, vp, sz, i32]
    lib.pvae_rollout_step.argtypes = [vp, i32, i32, i32, C.POINTER(vp), i64, vp, u64, u64, i32, f32, f32, f32, f32, vp, vp]
    lib.pvae_set_deterministic.argtypes = [vp, i32]
    lib.pvae_run_exchange.argtypes = [vp, vp]
    lib.pvae_set_exchange.argtypes = [vp, C.POINTER(u64), u64, i32, i32, i64, i64, i64, i32]
    lib.pvae_symm_allreduce.argtypes = [C.POINTER(u64), u64, i32, i32, i64, i64, i64, vp]
    lib.pvae_symm_flag_elems.restype = i64
    lib.pvae_symm_flag_elems.argtypes = []
    lib.pvae_gemm_bf16.argtypes = [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.pvae_debug_trace.argtypes = [vp, i32, i32]
    lib.pvae_launch_count.restype = u64
    lib.pvae_launch_count.argtypes = []
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("pvae_abi_version",):
            fn.restype = C.c_int
    if lib.pvae_abi_version() != PVAE_ABI_VERSION:
        raise PvaeError("libpvae_sm100.so ABI %d != binding ABI %d: rebuild" % (lib.pvae_abi_version(), PVAE_ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    """Turn a pvae_status into an exception carrying pvae_last_error()."""
    if rc == 0:
        return
    msg = load().pvae_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError("pvae: " + msg)
    raise PvaeError("pvae (status %d): %s" % (rc, msg))


def launch_count():
    return int(load().pvae_launch_count())
