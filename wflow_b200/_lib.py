"""ctypes binding of libwflowsbm.so (C-ABI: include/wflow_sbm.h).

There is deliberately no fallback: if the CUDA library has not been built, or no B200 is
visible, the error surfaces here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WFLOW_B200_LIB", os.path.join(_HERE, "libwflowsbm.so"))  # override: development aid

WF_MAX_LAYERS = 8


class WfConfig(C.Structure):
    _fields_ = [
        ("nrow", C.c_int32), ("ncol", C.c_int32),
        ("n_layer_thickness", C.c_int32),
        ("layer_thickness", C.c_double * WF_MAX_LAYERS),
        ("model_snow", C.c_int32), ("soil_inf_redu", C.c_int32), ("transfer_method", C.c_int32),
        ("ust", C.c_int32), ("glacier", C.c_int32), ("it_kin_l", C.c_int32), ("it_kin_r", C.c_int32),
        ("device", C.c_int32),
        ("timestepsecs", C.c_double), ("beta", C.c_double),
    ]


class WflowLibError(RuntimeError):
    pass


_lib = None

# every symbol include/wflow_sbm.h declares (checked by tests/test_cabi.py)
SYMBOLS = [
    "wf_last_error", "wf_version", "wf_create", "wf_destroy", "wf_num_cells", "wf_num_levels",
    "wf_num_river_cells", "wf_num_river_levels", "wf_get_network", "wf_catchment_total", "wf_set_field",
    "wf_set_field_uniform", "wf_get_field", "wf_request_output", "wf_num_fields", "wf_field_name",
    "wf_field_kind", "wf_field_device_ptr", "wf_cell_order", "wf_set_field_from_device", "wf_step",
    "wf_synchronize", "wf_last_step_ms", "wf_launch_count", "wf_stream", "wf_area_create", "wf_area_num_ids",
    "wf_area_ids", "wf_area_reduce", "wf_sample", "wf_quick_suspend", "wf_quick_resume",
    "wf_schedule_create", "wf_schedule_destroy", "wf_schedule_num_cells", "wf_schedule_num_levels",
    "wf_schedule_get", "wf_schedule_catchment_total", "wf_create_from_schedule", "wf_bind_field",
    "wf_timing_reset", "wf_timing_get", "wf_kinematic_wave_batch", "wf_kinematic_wave_ssf_batch",
    "wf_schedule_partition", "wf_lateral_plan", "wf_gauges_create", "wf_gauges_sample",
    "wf_estimate_iterations", "wf_set_substeps", "wf_pipeline_reserve", "wf_pipeline_set_forcing",
    "wf_pipeline_forcing_ptr", "wf_pipeline_capture", "wf_step_pipelined", "wf_gauges_sample_device", "wf_set_field_tiled", "wf_reservoirs_create", "wf_lakes_create",
    "wf_set_day_of_year", "wf_set_option", "wf_supply_iterations", "wf_irrigation_create", "wf_pipeline_mode", "wf_hbv_create", "wf_paddy_create",
]


def build():
    """Compile libwflowsbm.so in-tree (nvcc, sm_100a)."""
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise WflowLibError(
            "libwflowsbm.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C wflow_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    l.wf_last_error.restype = C.c_char_p
    l.wf_version.restype = C.c_char_p
    l.wf_field_name.restype = C.c_char_p
    l.wf_field_name.argtypes = [C.c_int]
    l.wf_field_kind.argtypes = [C.c_int]
    l.wf_create.argtypes = [C.POINTER(WfConfig), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    l.wf_destroy.argtypes = [C.c_void_p]
    l.wf_destroy.restype = None
    for name in ("wf_num_cells", "wf_num_levels", "wf_num_river_cells", "wf_num_river_levels", "wf_launch_count"):
        getattr(l, name).restype = C.c_int64
        getattr(l, name).argtypes = [C.c_void_p]
    l.wf_get_network.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_catchment_total.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_set_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    l.wf_set_field_from_device.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    l.wf_reservoirs_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_lakes_create.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] * 2
    l.wf_irrigation_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_paddy_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_pipeline_mode.argtypes = [C.c_void_p]
    l.wf_hbv_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_set_day_of_year.argtypes = [C.c_void_p, C.c_int]
    l.wf_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    l.wf_supply_iterations.argtypes = [C.c_void_p]
    l.wf_set_field_tiled.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int]
    l.wf_set_field_uniform.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
    l.wf_get_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_float]
    l.wf_request_output.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    l.wf_field_device_ptr.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    l.wf_cell_order.argtypes = [C.c_void_p, C.c_void_p]
    l.wf_step.argtypes = [C.c_void_p, C.c_int]
    l.wf_synchronize.argtypes = [C.c_void_p]
    l.wf_last_step_ms.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    l.wf_stream.argtypes = [C.c_void_p]
    l.wf_stream.restype = C.c_void_p
    l.wf_area_create.argtypes = [C.c_void_p, C.c_void_p]
    l.wf_area_num_ids.argtypes = [C.c_void_p, C.c_int]
    l.wf_area_ids.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    l.wf_area_reduce.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_void_p]
    l.wf_sample.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_float]
    l.wf_schedule_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    l.wf_schedule_destroy.argtypes = [C.c_void_p]
    l.wf_schedule_destroy.restype = None
    l.wf_schedule_num_cells.argtypes = [C.c_void_p]
    l.wf_schedule_num_cells.restype = C.c_int64
    l.wf_schedule_num_levels.argtypes = [C.c_void_p]
    l.wf_schedule_num_levels.restype = C.c_int64
    l.wf_schedule_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_schedule_catchment_total.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_create_from_schedule.argtypes = [C.POINTER(WfConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
    l.wf_bind_field.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    l.wf_timing_reset.argtypes = [C.c_void_p]
    l.wf_timing_get.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    l.wf_kinematic_wave_batch.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    l.wf_kinematic_wave_ssf_batch.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    l.wf_schedule_partition.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.wf_lateral_plan.argtypes = [C.c_void_p, C.c_void_p]
    l.wf_gauges_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    l.wf_gauges_sample.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_float, C.c_int]
    l.wf_gauges_sample_device.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_float]
    l.wf_estimate_iterations.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32)]
    l.wf_set_substeps.argtypes = [C.c_void_p, C.c_int, C.c_int]
    l.wf_pipeline_reserve.argtypes = [C.c_void_p, C.c_int]
    l.wf_pipeline_set_forcing.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
    l.wf_pipeline_forcing_ptr.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    l.wf_pipeline_capture.argtypes = [C.c_void_p, C.c_int]
    l.wf_step_pipelined.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_int]
    l.wf_quick_suspend.argtypes = [C.c_void_p]
    l.wf_quick_resume.argtypes = [C.c_void_p]
    _lib = l
    return l


def field_names():
    """Names of every field the library knows (forcing, statics, states, outputs on request)."""
    l = lib()
    return [l.wf_field_name(i).decode() for i in range(l.wf_num_fields())]


def check(rc):
    if rc < 0:
        raise WflowLibError("libwflowsbm: %s (code %d)" % (lib().wf_last_error().decode(), rc))
    return rc
