"""Fixture (de)serialisation shared by oracle/make_golden.py (writer, runs where the reference is)
and the tests (readers, run anywhere).  A problem dict (oracle/rp_oracle.py format) is flattened into
npz-friendly arrays plus one JSON string of scalars."""
import json

import numpy as np


def pack_problem(prob, prefix="p_"):
    out = {}
    meta = {k: prob[k] for k in ("x0_orientation", "x0_time_step", "lon_mode", "low_vel_mode", "dt", "N", "factor")}
    meta["draw_all"] = bool(prob.get("draw_all", False))
    meta["continuous"] = bool(prob.get("continuous", False))
    meta["constraints"] = list(prob["constraints"])
    meta["cost"] = prob["cost"]
    meta["vehicle"] = prob["vehicle"]
    meta["ccosy_limit"] = float(prob["ccosy"]["limit"])
    out[prefix + "meta"] = np.array(json.dumps(meta))
    for k in ("t", "lon", "d", "x0_lon", "x0_lat"):
        out[prefix + k] = np.asarray(prob[k], dtype=np.float64)
    for k, v in prob["ref"].items():
        out[prefix + "ref_" + k] = np.asarray(v, dtype=np.float64)
    out[prefix + "cc_path"] = np.asarray(prob["ccosy"]["path"], dtype=np.float64)
    out[prefix + "cc_S"] = np.asarray(prob["ccosy"]["S"], dtype=np.float64)
    out[prefix + "cc_normals"] = np.asarray(prob["ccosy"]["normals"], dtype=np.float64)
    out.update(pack_obstacles(prob["obstacles"], prefix + "ob_"))
    return out


def pack_obstacles(ob, prefix):
    out = {prefix + "static_boxes": np.asarray(ob["static_boxes"], dtype=np.float64).reshape(-1, 5),
           prefix + "dyn_t0": np.asarray(ob["dyn_t0"], dtype=np.int64),
           prefix + "dyn_lw": np.asarray(ob["dyn_lw"], dtype=np.float64).reshape(-1, 2),
           prefix + "boundary_boxes": np.asarray(ob["boundary_boxes"], dtype=np.float64).reshape(-1, 5),
           prefix + "boundary_tris": np.asarray(ob.get("boundary_tris", np.zeros((0, 6))), dtype=np.float64).reshape(-1, 6)}
    states = [np.asarray(s, dtype=np.float64).reshape(-1, 3) for s in ob["dyn_states"]]
    out[prefix + "dyn_len"] = np.array([len(s) for s in states], dtype=np.int64)
    out[prefix + "dyn_cat"] = np.concatenate(states, axis=0) if states else np.zeros((0, 3))
    return out


def unpack_obstacles(z, prefix):
    lens = z[prefix + "dyn_len"]
    cat = z[prefix + "dyn_cat"]
    states, off = [], 0
    for n in lens:
        states.append(cat[off:off + int(n)])
        off += int(n)
    return {"static_boxes": z[prefix + "static_boxes"], "dyn_t0": z[prefix + "dyn_t0"], "dyn_states": states,
            "dyn_lw": z[prefix + "dyn_lw"], "boundary_boxes": z[prefix + "boundary_boxes"],
            "boundary_tris": z[prefix + "boundary_tris"]}


def unpack_problem(z, prefix="p_"):
    meta = json.loads(str(z[prefix + "meta"]))
    prob = {k: meta[k] for k in ("x0_orientation", "x0_time_step", "lon_mode", "low_vel_mode", "dt", "N", "factor",
                                 "draw_all", "cost", "vehicle")}
    prob["constraints"] = tuple(meta["constraints"])
    prob["continuous"] = bool(meta.get("continuous", False))
    for k in ("t", "lon", "d", "x0_lon", "x0_lat"):
        prob[k] = z[prefix + k]
    prob["ref"] = {k: z[prefix + "ref_" + k] for k in ("ref_pos", "ref_theta", "ref_curv", "ref_curv_d")}
    prob["ccosy"] = {"path": z[prefix + "cc_path"], "S": z[prefix + "cc_S"], "normals": z[prefix + "cc_normals"],
                     "limit": meta["ccosy_limit"]}
    prob["obstacles"] = unpack_obstacles(z, prefix + "ob_")
    return prob
