"""ctypes binding of libparrot_b200.so (include/parrot_b200.h).

There is deliberately no CPU fallback: if the CUDA library is missing, or a
function is called without a CUDA device, this module raises.
"""
import ctypes as C
import os

from . import _build

_LIB = None


class ParrotConfig(C.Structure):
    """Mirror of ``parrot_config`` (include/parrot_b200.h)."""
    _fields_ = [
        ('input_dim', C.c_int32), ('output_dim', C.c_int32), ('rnn_h_dim', C.c_int32),
        ('readouts_dim', C.c_int32), ('weak_feedback', C.c_int32), ('full_feedback', C.c_int32),
        ('layer_norm', C.c_int32), ('use_speaker', C.c_int32), ('num_speakers', C.c_int32),
        ('speaker_dim', C.c_int32), ('which_cost', C.c_int32), ('k_gmm', C.c_int32),
        ('num_characters', C.c_int32), ('attention_type', C.c_int32), ('attention_size', C.c_int32),
        ('encoder_type', C.c_int32), ('encoder_dim', C.c_int32), ('encoder_time_axis', C.c_int32),
        ('sampling_bias', C.c_float), ('epsilon', C.c_float), ('attention_alignment', C.c_float),
        ('sharpening_coeff', C.c_float), ('timing_coeff', C.c_float),
        ('batch_size', C.c_int32), ('seq_len', C.c_int32), ('text_len', C.c_int32),
        ('gemm_impl', C.c_int32), ('sampling', C.c_int32),
    ]


# every symbol include/parrot_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    'parrot_last_error', 'parrot_abi_version', 'parrot_param_count', 'parrot_param_info',
    'parrot_workspace_bytes', 'parrot_create', 'parrot_destroy', 'parrot_buffer_info',
    'parrot_pack_weights', 'parrot_mark_params_dirty', 'parrot_set_profiling', 'parrot_get_profile', 'parrot_debug_time_table', 'parrot_debug_set_stamps', 'parrot_encoder_fwd', 'parrot_encoder_bwd',
    'parrot_decoder_scan_fwd', 'parrot_decoder_scan_bwd', 'parrot_readout_emit_fwd',
    'parrot_readout_emit_bwd', 'parrot_attention_step', 'parrot_compute_cost', 'parrot_backward',
    'parrot_sample_scan', 'parrot_adam_clip_step', 'parrot_gemm_nt',
    'parrot_gemm_nt_workspace_bytes', 'parrot_launch_count',
    'parrot_comm_unique_id', 'parrot_comm_init', 'parrot_comm_allreduce', 'parrot_comm_info', 'parrot_comm_destroy',
]


def lib_path():
    return _build.LIB


def load():
    """Load (building first if sources are newer and nvcc is present) the CUDA library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get('PARROT_B200_LIB')     # experiment hook: a variant build of the same sources
    if not path:
        path = _build.LIB
        try:
            _build.build()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise RuntimeError(
            'parrot_b200: %s not found. Run `python -c "import __graft_entry__ as g; g.build()"`; '
            'there is no CPU fallback.' % path)
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.parrot_last_error.restype = C.c_char_p
    lib.parrot_launch_count.restype = C.c_int64
    lib.parrot_gemm_nt_workspace_bytes.restype = C.c_size_t
    lib.parrot_gemm_nt_workspace_bytes.argtypes = [C.c_int32] * 3
    lib.parrot_param_count.argtypes = [C.POINTER(ParrotConfig), C.POINTER(C.c_int32), C.POINTER(C.c_int64)]
    lib.parrot_param_info.argtypes = [C.POINTER(ParrotConfig), C.c_int32, C.c_char_p, C.c_int32,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.parrot_workspace_bytes.argtypes = [C.POINTER(ParrotConfig), C.POINTER(C.c_size_t)]
    lib.parrot_create.argtypes = [C.POINTER(ParrotConfig), vp, vp, vp, C.c_size_t, vp, C.POINTER(vp)]
    lib.parrot_destroy.argtypes = [vp]
    lib.parrot_buffer_info.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.parrot_pack_weights.argtypes = [vp, vp]
    lib.parrot_mark_params_dirty.argtypes = [vp]
    lib.parrot_encoder_fwd.argtypes = [vp, vp, vp, vp]
    lib.parrot_encoder_bwd.argtypes = [vp, vp]
    lib.parrot_decoder_scan_fwd.argtypes = [vp, vp, vp, C.c_float, C.c_float, vp]
    lib.parrot_decoder_scan_bwd.argtypes = [vp, vp]
    lib.parrot_readout_emit_fwd.argtypes = [vp, vp, vp, vp, vp]
    lib.parrot_readout_emit_bwd.argtypes = [vp, C.c_int, vp]
    lib.parrot_attention_step.argtypes = [C.POINTER(ParrotConfig)] + [vp] * 10 + [C.c_int, vp]
    lib.parrot_compute_cost.argtypes = [vp, vp, vp, vp, vp, vp, C.c_float, vp, C.c_float, vp, vp, vp, vp]
    lib.parrot_backward.argtypes = [vp, C.c_int, vp]
    lib.parrot_sample_scan.argtypes = [vp, vp, vp, vp, vp, vp, C.c_uint64, vp]
    lib.parrot_adam_clip_step.argtypes = [vp, vp, vp, vp, C.c_int64, C.c_float, vp, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_float, C.c_int64, vp, vp, vp]
    lib.parrot_debug_time_table.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), vp, vp]
    lib.parrot_debug_set_stamps.argtypes = [vp, vp, C.c_int]
    lib.parrot_set_profiling.argtypes = [vp, C.c_int]
    lib.parrot_get_profile.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.parrot_gemm_nt.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, C.c_size_t, vp]
    lib.parrot_comm_unique_id.argtypes = [vp]
    lib.parrot_comm_init.argtypes = [C.c_int32, C.c_int32, vp, C.POINTER(vp)]
    lib.parrot_comm_allreduce.argtypes = [vp, vp, C.c_int64, vp]
    lib.parrot_comm_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.parrot_comm_destroy.argtypes = [vp]
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError(load().parrot_last_error().decode())


def param_layout(cfg):
    """[(name, offset, shape)] and total floats, straight from the library."""
    lib = load()
    n = C.c_int32()
    total = C.c_int64()
    check(lib.parrot_param_count(C.byref(cfg), C.byref(n), C.byref(total)))
    out = []
    buf = C.create_string_buffer(256)
    for i in range(n.value):
        off = C.c_int64(); r = C.c_int32(); c = C.c_int32()
        check(lib.parrot_param_info(C.byref(cfg), i, buf, 256, C.byref(off), C.byref(r), C.byref(c)))
        shape = (r.value, c.value) if c.value else (r.value,)
        out.append((buf.value.decode(), off.value, shape))
    return out, total.value
