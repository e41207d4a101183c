/*
 * rp_pack.c -- CPython helper for the OUTPUT PACKING of ReactivePlanner.plan()
 * (reference commonroad_rp/reactive_planner.py:514-568, _compute_trajectory_pair): the winner's 14 x (N + 1) state block
 * becomes N + 1 planner-state objects plus the two curvilinear state lists.  In Python that is 42+ small objects and a
 * dozen numpy calls per replanning cycle -- a quarter of plan()'s wall time once the device work takes 40 us.
 *
 * Host glue only: no planner arithmetic beyond what :520-556 does per state (steering angle atan2(wheelbase * kappa, 1),
 * yaw rate (theta[i] - theta[i-1]) / dt, orientation folded by whole turns into [lo, hi] as utility/general.py:49-55).
 * Used only with the package's own stand-in state classes (no commonroad-io); reactive_planner.py falls back to the
 * Python loop when this module is not built.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <math.h>

static PyObject *k_time_step, *k_position, *k_steering_angle, *k_velocity, *k_orientation, *k_acceleration, *k_yaw_rate;
static PyObject* k_dict;

static const double TWO_PI = 6.283185307179586;

/* new instance of a plain Python class with the given instance dictionary (cls.__new__(cls); obj.__dict__ = d) */
static PyObject* instance_with_dict(PyTypeObject* cls, PyObject* empty, PyObject* d) {
    PyObject* obj = cls->tp_new(cls, empty, NULL);
    if (!obj) return NULL;
    if (PyObject_SetAttr(obj, k_dict, d) < 0) {
        Py_DECREF(obj);
        return NULL;
    }
    return obj;
}

static int set_steal(PyObject* d, PyObject* key, PyObject* value) {
    if (!value) return -1;
    const int rc = PyDict_SetItem(d, key, value);
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
        pos_fast = PySequence_Fast(positions, "positions must be a sequence");
        if (!pos_fast) goto done;
        if (PySequence_Fast_GET_SIZE(pos_fast) != n) {
            PyErr_SetString(PyExc_ValueError, "positions must have one entry per state");
            goto done;
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
            PyObject* dct = PyDict_New();
            if (!dct) goto done;
            PyObject* pos = PySequence_Fast_GET_ITEM(pos_fast, i);
            if (set_steal(dct, k_time_step, PyLong_FromLong(ts)) < 0 || PyDict_SetItem(dct, k_position, pos) < 0 ||
                set_steal(dct, k_steering_angle, PyFloat_FromDouble(steering)) < 0 ||
                set_steal(dct, k_velocity, PyFloat_FromDouble(v[i])) < 0 ||
                set_steal(dct, k_orientation, PyFloat_FromDouble(theta)) < 0 ||
                set_steal(dct, k_acceleration, PyFloat_FromDouble(a[i])) < 0 ||
                set_steal(dct, k_yaw_rate, PyFloat_FromDouble(yaw_rate)) < 0) {
                Py_DECREF(dct);
                goto done;
            }
            PyObject* st = instance_with_dict(cls, empty, dct);
            Py_DECREF(dct);
            if (!st) goto done;
            PyList_SET_ITEM(cart, i, st);
            PyObject* l3 = Py_BuildValue("[ddd]", s[i], sv[i], sa[i]);
            PyObject* t3 = Py_BuildValue("[ddd]", d[i], dv[i], da[i]);
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

static PyMethodDef methods[] = {
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
    k_dict = PyUnicode_InternFromString("__dict__");
    return PyModule_Create(&module);
}
