"""ctypes declarations of the C ABI in include/pkb200.h (pure declarations).

``declare(cdll)`` attaches argument/return types for every entry point to a
loaded library object and returns it.  ``SYMBOLS`` lists every function the
header declares; tests check that the built library exports all of them.
"""
import ctypes as C

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_ll_p = C.POINTER(C.c_longlong)

PKB_OK = 0
PKB_EINVAL, PKB_ECUDA, PKB_ENOMEM, PKB_ELIMIT, PKB_ESTATE = -1, -2, -3, -4, -5

ST_HPROB_RANGE = 1
ST_NEG_LOSS = 2
ST_PMF_NEG = 4
ST_PMF_GT1 = 8
ST_PMF_NEG2 = 16
ST_TOT_GT1 = 32
ST_WARNED = 64
ST_BORDERLINE = 128


class DayArgs(C.Structure):
    _fields_ = [('hparams', C.c_double * 7), ('dparams', C.c_double * 3), ('dlparams', C.c_double * 3),
                ('mu_r', C.c_double), ('rad_dist', C.c_double), ('start_time', C.c_double),
                ('n_periods', C.c_int), ('rad_res', C.c_int), ('wind_day', C.c_int), ('single', C.c_int),
                ('kind', C.c_int), ('pad_', C.c_int), ('sprd_factor', C.c_double), ('sprd_drift', C.c_double * 2)]


class DayMeta(C.Structure):
    _fields_ = [('loss', C.c_double), ('pmfsum', C.c_double), ('total', C.c_double), ('kept_sum', C.c_double),
                ('add', C.c_double), ('rad', C.c_int), ('nnz', C.c_int), ('status', C.c_int), ('ext', C.c_int),
                ('hl', C.c_int), ('pad_', C.c_int)]


class StepMeta(C.Structure):
    _fields_ = [('padmax', C.c_double), ('ksum', C.c_double), ('add', C.c_double), ('padabs', C.c_double),
                ('kcnt', C.c_longlong), ('flag', C.c_int), ('spec', C.c_int),
                ('wr0', C.c_int), ('wr1', C.c_int), ('wc0', C.c_int), ('wc1', C.c_int),
                ('er0', C.c_int), ('er1', C.c_int), ('ec0', C.c_int), ('ec1', C.c_int)]


class SolveArgs(C.Structure):
    _fields_ = [('wind', C.c_void_p), ('wind_on_device', C.c_int), ('nd_wind', C.c_int), ('periods', C.c_int),
                ('ndays', C.c_int), ('day', DayArgs), ('prob_model', C.c_int), ('r_dur', C.c_int),
                ('r_number', C.c_double), ('r_dist', c_double_p), ('r_start', C.c_double), ('negval', C.c_double),
                ('want_dense_host', C.c_int), ('want_coo', C.c_int), ('keep_dense_device', C.c_int),
                ('sprd', C.c_int), ('sprd_factor', C.c_double), ('sprd_drift', C.c_double * 2), ('sprd_factors', c_double_p),
                ('out_on_device', C.c_int), ('keep_pre_device', C.c_int)]


class Projection(C.Structure):
    _fields_ = [('nsets', C.c_int), ('set_ptr', c_int_p), ('set_cells', c_int_p), ('nrows', C.c_int), ('row_ptr', c_int_p),
                ('ngroups', C.c_int), ('grp_ptr', c_int_p), ('term_day', c_int_p), ('term_set', c_int_p), ('term_w', c_double_p)]


_H = C.c_void_p        # opaque handles
_HP = C.POINTER(C.c_void_p)

_SIGS = {
    'pkb_last_error': (C.c_char_p, []),
    'pkb_version': (C.c_int, []),
    'pkb_create': (C.c_int, [C.c_int, _HP]),
    'pkb_destroy': (C.c_int, [_H]),
    'pkb_sync': (C.c_int, [_H]),
    'pkb_set_option': (C.c_int, [_H, C.c_char_p, C.c_double]),
    'pkb_timing': (C.c_int, [_H, c_double_p]),
    'pkb_launch_count': (C.c_longlong, [_H]),
    'pkb_mark': (C.c_int, [_H, C.c_int]),
    'pkb_elapsed_ms': (C.c_int, [_H, C.c_int, C.c_int, c_double_p]),
    'pkb_profile_enable': (C.c_int, [_H, C.c_int]),
    'pkb_profile_reset': (C.c_int, [_H]),
    'pkb_profile_get': (C.c_int, [_H, C.c_char_p, c_ll_p, c_double_p]),
    'pkb_wind_interp': (C.c_int, [_H, c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, c_double_p]),
    'pkb_hprob': (C.c_int, [_H, c_double_p, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    'pkb_mvn_cdf': (C.c_int, [_H, C.c_double, c_double_p, c_double_p, c_double_p, C.c_int, c_int_p]),
    'pkb_kernels_build': (C.c_int, [_H, c_double_p, C.c_int, C.c_int, C.POINTER(DayArgs), C.c_int, C.c_int, _HP]),
    'pkb_kset_meta': (C.c_int, [_H, C.c_int, C.POINTER(DayMeta)]),
    'pkb_kset_get': (C.c_int, [_H, C.c_int, c_double_p]),
    'pkb_kset_get_pre': (C.c_int, [_H, C.c_int, c_double_p]),
    'pkb_kset_racc': (C.c_int, [_H]),
    'pkb_kset_periods': (C.c_int, [_H, C.c_int, c_int_p, c_double_p]),
    'pkb_kset_destroy': (C.c_int, [_H]),
    'pkb_chain_create': (C.c_int, [_H, C.c_int, C.c_int, _HP]),
    'pkb_chain_destroy': (C.c_int, [_H]),
    'pkb_chain_set_state': (C.c_int, [_H, c_double_p]),
    'pkb_chain_set_state_kernel': (C.c_int, [_H, _H, C.c_int]),
    'pkb_chain_conv': (C.c_int, [_H, c_double_p, C.c_int]),
    'pkb_chain_conv_kernel': (C.c_int, [_H, _H, C.c_int]),
    'pkb_chain_get_cursol': (C.c_int, [_H, C.c_double, C.c_int, C.c_int, c_double_p, C.POINTER(StepMeta)]),
    'pkb_chain_back_solve': (C.c_int, [_H, C.POINTER(c_double_p), c_int_p, C.c_int, C.c_double, c_double_p, c_int_p]),
    'pkb_chain_population': (C.c_int, [_H, C.c_int, c_double_p, C.c_double, C.c_double, C.c_int, C.c_double, C.c_int,
                                       c_double_p, c_double_p]),
    'pkb_chain_get_state': (C.c_int, [_H, c_double_p]),
    'pkb_chain_dims': (C.c_int, [_H, c_int_p, c_int_p, c_int_p]),
    'pkb_debug_fft': (C.c_int, [_H, C.c_int, c_double_p, c_double_p, C.c_int]),
    'pkb_smooth_len': (C.c_int, [C.c_int]),
    'pkb_solve': (C.c_int, [_H, C.POINTER(SolveArgs), _HP]),
    'pkb_solve_batch': (C.c_int, [_H, C.POINTER(SolveArgs), c_double_p, C.c_int, c_int_p, C.c_int, c_double_p, c_int_p]),
    'pkb_project': (C.c_int, [_H, C.POINTER(Projection), c_double_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    'pkb_solve_batch_projected': (C.c_int, [_H, C.POINTER(SolveArgs), c_double_p, C.c_int, c_int_p, C.c_int, C.POINTER(Projection),
                                            c_double_p, c_int_p]),
    'pkb_result_project': (C.c_int, [_H, c_int_p, C.c_int, C.POINTER(Projection), c_double_p]),
    'pkb_set_stream': (C.c_int, [_H, C.c_void_p]),
    'pkb_kset_export_device': (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int]),
    'pkb_kset_from_device': (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int, c_int_p, C.c_int, _HP]),
    'pkb_dist_plan': (C.c_int, [_H, C.c_int, C.c_int, C.c_int, c_ll_p, c_int_p, c_int_p, c_int_p]),
    'pkb_dist_create': (C.c_int, [_H, _H, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _HP]),
    'pkb_dist_step_cols': (C.c_int, [_H, C.c_int]),
    'pkb_dist_step_rows': (C.c_int, [_H]),
    'pkb_dist_step_emit': (C.c_int, [_H, C.c_int]),
    'pkb_dist_finish': (C.c_int, [_H, c_double_p, c_int_p]),
    'pkb_dist_destroy': (C.c_int, [_H]),
    'pkb_result_info': (C.c_int, [_H, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    'pkb_result_window_steps': (C.c_int, [_H, c_int_p]),
    'pkb_result_day_meta': (C.c_int, [_H, C.c_int, C.POINTER(DayMeta), C.POINTER(StepMeta)]),
    'pkb_result_cohort_meta': (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(StepMeta)]),
    'pkb_result_dense': (C.c_int, [_H, C.c_int, c_double_p]),
    'pkb_result_pre': (C.c_int, [_H, C.c_int, c_double_p]),
    'pkb_result_coo': (C.c_int, [_H, C.POINTER(c_ll_p), C.POINTER(c_int_p), C.POINTER(c_int_p), C.POINTER(c_double_p)]),
    'pkb_result_csr': (C.c_int, [_H, C.POINTER(c_ll_p), C.POINTER(c_ll_p), C.POINTER(c_int_p), C.POINTER(c_double_p)]),
    'pkb_result_sample': (C.c_int, [_H, c_int_p, C.c_int, c_double_p]),
    'pkb_result_device_ptr': (C.c_int, [_H, _HP]),
    'pkb_result_destroy': (C.c_int, [_H]),
}

SYMBOLS = tuple(sorted(_SIGS))


def declare(cdll):
    """Attach restype/argtypes for every entry point; raises AttributeError if
    the library does not export one of them."""
    for name, (res, args) in _SIGS.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll
