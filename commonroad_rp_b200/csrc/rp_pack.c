/*
 * rp_pack.c -- CPython helpers around the C-ABI for ReactivePlanner.plan(): the output packing and the rp_plan_levels call.
 *
 * (1) OUTPUT PACKING of ReactivePlanner.plan()
 * (reference commonroad_rp/reactive_planner.py:514-568, _compute_trajectory_pair): the winner's 14 x (N + 1) state block
 * becomes N + 1 planner-state objects plus the two curvilinear state lists.  In Python that is 42+ small objects and a
 * dozen numpy calls per replanning cycle -- a quarter of plan()'s wall time once the device work takes 40 us.
 *
 * Host glue only: no planner arithmetic beyond what :520-556 does per state (steering angle atan2(wheelbase * kappa, 1),
 * yaw rate (theta[i] - theta[i-1]) / dt, orientation folded by whole turns into [lo, hi] as utility/general.py:49-55).
 * Used only with the package's own stand-in state classes (no commonroad-io); reactive_planner.py falls back to the
 * Python loop when this module is not built.
 *
 * (2) plan_levels(): the per-cycle call of rp_plan_levels (include/rp_b200.h) without ctypes -- the levels' sample arrays
 * are concatenated on the stack and the library is entered through its exported address.  Same C-ABI, same arguments as
 * _lib.Engine.plan_levels' ctypes path (which stays as the fallback); 6 us less host time per replanning cycle.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>
#ifdef RP_PACK_NUMPY          /* build.py defines it when the numpy headers are there: positions are created here */
#define NPY_NO_DEPRECATED_API NPY_1_7_API_VERSION
#include <numpy/arrayobject.h>
#endif

static PyObject *k_time_step, *k_position, *k_steering_angle, *k_velocity, *k_orientation, *k_acceleration, *k_yaw_rate;

static const double TWO_PI = 6.283185307179586;

/* [a, b, c] as a new list of floats */
static PyObject* list3(double a, double b, double c) {
    PyObject* l = PyList_New(3);
    if (!l) return NULL;
    PyObject *x = PyFloat_FromDouble(a), *y = PyFloat_FromDouble(b), *z = PyFloat_FromDouble(c);
    if (!x || !y || !z) {
        Py_XDECREF(x); Py_XDECREF(y); Py_XDECREF(z);
        Py_DECREF(l);
        return NULL;
    }
    PyList_SET_ITEM(l, 0, x);
    PyList_SET_ITEM(l, 1, y);
    PyList_SET_ITEM(l, 2, z);
    return l;
}

static int attr_steal(PyObject* obj, PyObject* key, PyObject* value) {
    if (!value) return -1;
    const int rc = PyObject_SetAttr(obj, key, value);
    Py_DECREF(value);
    return rc;
}

/* pack(state_cls, block, positions, t0, factor, dt, wheelbase, yaw_rate_0, lo, hi) -> (cart_list, lon_list, lat_list)
 *   block      C-contiguous float64 buffer [14][n]  (rp_state_row order: x y theta v a kappa kappa_dot s d theta_cl
 *              s_dot s_ddot d_dot d_ddot)
 *   positions  sequence of n position objects (2-vectors) for the Cartesian states */
static PyObject* pack(PyObject* self, PyObject* args) {
    PyObject *cls_obj, *block_obj, *positions;
    long t0, factor;
    double dt, wheelbase, yaw0, lo, hi;
    if (!PyArg_ParseTuple(args, "OOOllddddd", &cls_obj, &block_obj, &positions, &t0, &factor, &dt, &wheelbase, &yaw0, &lo, &hi))
        return NULL;
    if (!PyType_Check(cls_obj)) {
        PyErr_SetString(PyExc_TypeError, "state class expected");
        return NULL;
    }
    PyTypeObject* cls = (PyTypeObject*)cls_obj;
    Py_buffer view;
    if (PyObject_GetBuffer(block_obj, &view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) return NULL;
    PyObject *cart = NULL, *lon = NULL, *lat = NULL, *empty = NULL, *result = NULL, *pos_fast = NULL;
    if (view.itemsize != 8 || view.len % (14 * 8) != 0 || (view.format && view.format[0] != 'd')) {
        PyErr_SetString(PyExc_ValueError, "block must be a C-contiguous float64 array of shape (14, n)");
        goto done;
    }
    {
        const Py_ssize_t n = view.len / (14 * 8);
        const double* b = (const double*)view.buf;
        const double *th = b + 2 * n, *v = b + 3 * n, *a = b + 4 * n, *kap = b + 5 * n;
        const double *s = b + 7 * n, *d = b + 8 * n, *sv = b + 10 * n, *sa = b + 11 * n, *dv = b + 12 * n, *da = b + 13 * n;
        const int make_positions = positions == Py_None;      /* None: one new (2,) float64 array per state, made here */
#ifndef RP_PACK_NUMPY
        if (make_positions) {
            PyErr_SetString(PyExc_TypeError, "positions=None needs the numpy build of _rp_pack");
            goto done;
        }
#endif
        if (!make_positions) {
            pos_fast = PySequence_Fast(positions, "positions must be a sequence");
            if (!pos_fast) goto done;
            if (PySequence_Fast_GET_SIZE(pos_fast) != n) {
                PyErr_SetString(PyExc_ValueError, "positions must have one entry per state");
                goto done;
            }
        }
        empty = PyTuple_New(0);
        cart = PyList_New(n);
        lon = PyList_New(n);
        lat = PyList_New(n);
        if (!empty || !cart || !lon || !lat) goto done;
        long ts = t0;
        for (Py_ssize_t i = 0; i < n; ++i, ts += factor) {
            double theta = th[i];
            while (theta < lo) theta += TWO_PI;               /* shift_orientation (utility/general.py:49-55) */
            while (theta > hi) theta -= TWO_PI;
            const double yaw_rate = i == 0 ? yaw0 : (th[i] - th[i - 1]) / dt;     /* :536-539 */
            const double steering = atan2(wheelbase * kap[i], 1.0);              /* :540-541 */
            /* a fresh instance whose attributes are set one by one (the interpreter's own fast path for instance
               attributes; same attribute order as ReactivePlannerState.__init__) */
            PyObject* st = cls->tp_new(cls, empty, NULL);
            if (!st) goto done;
            PyObject* pos = NULL;
            int pos_owned = 0;
#ifdef RP_PACK_NUMPY
            if (make_positions) {
                npy_intp two = 2;
                pos = PyArray_SimpleNew(1, &two, NPY_DOUBLE);
                if (!pos) {
                    Py_DECREF(st);
                    goto done;
                }
                double* pd = (double*)PyArray_DATA((PyArrayObject*)pos);
                pd[0] = b[i];
                pd[1] = b[n + i];
                pos_owned = 1;
            }
#endif
            if (!pos) pos = PySequence_Fast_GET_ITEM(pos_fast, i);
            int rc = attr_steal(st, k_time_step, PyLong_FromLong(ts));
            if (rc == 0) rc = PyObject_SetAttr(st, k_position, pos);
            if (pos_owned) Py_DECREF(pos);
            if (rc < 0 || attr_steal(st, k_steering_angle, PyFloat_FromDouble(steering)) < 0 ||
                attr_steal(st, k_velocity, PyFloat_FromDouble(v[i])) < 0 ||
                attr_steal(st, k_orientation, PyFloat_FromDouble(theta)) < 0 ||
                attr_steal(st, k_acceleration, PyFloat_FromDouble(a[i])) < 0 ||
                attr_steal(st, k_yaw_rate, PyFloat_FromDouble(yaw_rate)) < 0) {
                Py_DECREF(st);
                goto done;
            }
            PyList_SET_ITEM(cart, i, st);
            PyObject* l3 = list3(s[i], sv[i], sa[i]);
            PyObject* t3 = list3(d[i], dv[i], da[i]);
            if (!l3 || !t3) {
                Py_XDECREF(l3);
                Py_XDECREF(t3);
                goto done;
            }
            PyList_SET_ITEM(lon, i, l3);
            PyList_SET_ITEM(lat, i, t3);
        }
        result = PyTuple_Pack(3, cart, lon, lat);
    }
done:
    PyBuffer_Release(&view);
    Py_XDECREF(pos_fast);
    Py_XDECREF(empty);
    Py_XDECREF(cart);
    Py_XDECREF(lon);
    Py_XDECREF(lat);
    return result;
}

/* ---- rp_plan_levels without ctypes -------------------------------------------------------------------------------- */
#include <stdint.h>
#include <string.h>

typedef int (*rp_plan_levels_fn)(void* ctx, const void* in, int n_levels, const int32_t* n_t, const int32_t* n_lon,
                                 const int32_t* n_d, const double* t_cat, const int32_t* traj_len_cat, const double* lon_cat,
                                 const double* d_cat, void* out, int32_t* n_evaluated, int32_t* chosen);

#define RP_MAX_LEVELS 4
#define RP_MAX_SAMPLES 448

/* copies a 1-D contiguous buffer of `fmt` items into dst[*off ..]; returns the item count or -1 (exception set) */
static Py_ssize_t take(PyObject* obj, char fmt, Py_ssize_t itemsize, void* dst, Py_ssize_t* off) {
    Py_buffer v;
    if (PyObject_GetBuffer(obj, &v, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) < 0) return -1;
    const char* f = v.format ? v.format : "B";
    if (*f == '=' || *f == '<' || *f == '@') ++f;
    Py_ssize_t n = -1;
    if (v.itemsize != itemsize || f[0] != fmt || f[1] != 0 || v.ndim != 1)
        PyErr_SetString(PyExc_TypeError, "sample arrays: 1-D float64 (t, lon, d) and int32 (traj_len)");
    else if (*off + v.len / itemsize > RP_MAX_SAMPLES)
        PyErr_SetString(PyExc_ValueError, "too many samples for one launch");
    else {
        n = v.len / itemsize;
        memcpy((char*)dst + *off * itemsize, v.buf, (size_t)v.len);
        *off += n;
    }
    PyBuffer_Release(&v);
    return n;
}

/* plan_levels(fn_addr, ctx_addr, inputs_addr, levels, out_addr) -> (rc, chosen, n_evaluated, (count per level ...))
 *   levels   sequence of (t, lon, d, traj_len) per sampling level, escalation order
 *   out_addr rp_plan_result[4] owned by the caller */
static PyObject* plan_levels(PyObject* self, PyObject* args) {
    unsigned long long fn_addr, ctx_addr, in_addr, out_addr;
    PyObject* levels;
    if (!PyArg_ParseTuple(args, "KKKOK", &fn_addr, &ctx_addr, &in_addr, &levels, &out_addr)) return NULL;
    PyObject* seq = PySequence_Fast(levels, "levels must be a sequence");
    if (!seq) return NULL;
    const Py_ssize_t n_levels = PySequence_Fast_GET_SIZE(seq);
    double t_cat[RP_MAX_SAMPLES], lon_cat[RP_MAX_SAMPLES], d_cat[RP_MAX_SAMPLES];
    int32_t tl_cat[RP_MAX_SAMPLES], n_t[RP_MAX_LEVELS], n_lon[RP_MAX_LEVELS], n_d[RP_MAX_LEVELS];
    Py_ssize_t ot = 0, otl = 0, ol = 0, od = 0;
    PyObject* result = NULL;
    if (n_levels < 1 || n_levels > RP_MAX_LEVELS) {
        PyErr_SetString(PyExc_ValueError, "1 .. 4 levels");
        goto done;
    }
    for (Py_ssize_t j = 0; j < n_levels; ++j) {
        PyObject* lv = PySequence_Fast_GET_ITEM(seq, j);
        if (!PyTuple_Check(lv) || PyTuple_GET_SIZE(lv) != 4) {
            PyErr_SetString(PyExc_TypeError, "level = (t, lon, d, traj_len)");
            goto done;
        }
        const Py_ssize_t a = take(PyTuple_GET_ITEM(lv, 0), 'd', 8, t_cat, &ot);
        const Py_ssize_t b = a < 0 ? -1 : take(PyTuple_GET_ITEM(lv, 1), 'd', 8, lon_cat, &ol);
        const Py_ssize_t c = b < 0 ? -1 : take(PyTuple_GET_ITEM(lv, 2), 'd', 8, d_cat, &od);
        const Py_ssize_t e = c < 0 ? -1 : take(PyTuple_GET_ITEM(lv, 3), 'i', 4, tl_cat, &otl);
        if (e < 0) goto done;
        if (e != a) {
            PyErr_SetString(PyExc_ValueError, "one traj_len per sampled horizon");
            goto done;
        }
        n_t[j] = (int32_t)a; n_lon[j] = (int32_t)b; n_d[j] = (int32_t)c;
    }
    {
        int32_t n_eval = 0, chosen = 0;
        const int rc = ((rp_plan_levels_fn)(uintptr_t)fn_addr)((void*)(uintptr_t)ctx_addr, (const void*)(uintptr_t)in_addr,
                                                              (int)n_levels, n_t, n_lon, n_d, t_cat, tl_cat, lon_cat, d_cat,
                                                              (void*)(uintptr_t)out_addr, &n_eval, &chosen);
        PyObject* counts = PyTuple_New(n_levels);
        if (!counts) goto done;
        for (Py_ssize_t j = 0; j < n_levels; ++j)
            PyTuple_SET_ITEM(counts, j, PyLong_FromLongLong((long long)n_t[j] * n_lon[j] * n_d[j]));
        result = Py_BuildValue("(iiiN)", rc, (int)chosen, (int)n_eval, counts);
    }
done:
    Py_DECREF(seq);
    return result;
}

static PyMethodDef methods[] = {
    {"plan_levels", plan_levels, METH_VARARGS, "rp_plan_levels through its exported address (no ctypes marshalling)"},
    {"pack", pack, METH_VARARGS, "winner state block -> (Cartesian state list, lon list, lat list)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_rp_pack", "output packing of ReactivePlanner.plan()", -1, methods};

PyMODINIT_FUNC PyInit__rp_pack(void) {
    k_time_step = PyUnicode_InternFromString("time_step");
    k_position = PyUnicode_InternFromString("position");
    k_steering_angle = PyUnicode_InternFromString("steering_angle");
    k_velocity = PyUnicode_InternFromString("velocity");
    k_orientation = PyUnicode_InternFromString("orientation");
    k_acceleration = PyUnicode_InternFromString("acceleration");
    k_yaw_rate = PyUnicode_InternFromString("yaw_rate");
    PyObject* m = PyModule_Create(&module);
    if (!m) return NULL;
#ifdef RP_PACK_NUMPY
    import_array();
    PyModule_AddIntConstant(m, "MAKES_POSITIONS", 1);
#else
    PyModule_AddIntConstant(m, "MAKES_POSITIONS", 0);
#endif
    return m;
}
