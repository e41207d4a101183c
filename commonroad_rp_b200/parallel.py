"""Multi-GPU plumbing: one process per GPU, bundles sharded t-major over the ranks, one small
NCCL exchange for the global arg-min.  The reference's only parallelism is
fork + pickle of candidate chunks (reactive_planner.py:1084-1111); candidates are independent up to
the final arg-min, so the only data that ever crosses NVLink is a 32-byte record per rank.
"""
import torch
import torch.distributed as dist


def shard_range(n_candidates: int, rank: int, world: int):
    """Contiguous (t-major) tile of the enumeration space owned by ``rank``: (first, count)."""
    per = (n_candidates + world - 1) // world
    first = min(rank * per, n_candidates)
    return first, max(0, min(per, n_candidates - first))


def merge_records(gathered: torch.Tensor):
    """Lexicographic min on (cost, enumeration index) over the ranks' records [world, 4]; returns
    (winner[2] = [cost, index] (+inf = none), totals[2] = [n_infeasible_kinematics, n_feasible])."""
    cost = gathered[:, 0]
    idx = gathered[:, 1]
    cmin = cost.min()
    imin = torch.where(cost == cmin, idx, torch.full_like(idx, float("inf"))).min()
    return torch.stack([cmin, imin]), gathered[:, 2:4].sum(dim=0)


def global_argmin(engine, rec: torch.Tensor, world: int, group=None):
    """After ``engine.grid_launch()`` on every rank: all-gather the shard records, pick the global
    winner and count the colliders ranked before it.  Everything stays on the device and on the
    current stream; returns device tensors (winner[2], totals[2], n_collision_before[1]).
    Per cycle: record export, one all-gather of 32 bytes per rank, a one-warp merge kernel, the shard's count of
    colliders before the global winner, one scalar all-reduce -- workspace tensors are allocated once per engine."""
    ws = getattr(engine, "_argmin_ws", None)
    if ws is None or ws[0].numel() != world * rec.numel() or ws[0].device != rec.device:
        ws = (torch.empty(world * rec.numel(), dtype=rec.dtype, device=rec.device),
              torch.empty(2, dtype=torch.float64, device=rec.device), torch.empty(2, dtype=torch.float64, device=rec.device),
              torch.empty(1, dtype=torch.float64, device=rec.device))
        engine._argmin_ws = ws
    gathered, winner, totals, before = ws
    engine.export_record_dev(rec.data_ptr())
    dist.all_gather_into_tensor(gathered, rec, group=group)
    if rec.is_cuda:
        engine.merge_records_dev(gathered.data_ptr(), world, winner.data_ptr(), totals.data_ptr())
    else:
        w, t = merge_records(gathered.view(world, rec.numel()))
        winner.copy_(w)
        totals.copy_(t)
    engine.count_colliders_before_dev(winner.data_ptr(), before.data_ptr())
    dist.all_reduce(before, group=group)
    return winner, totals, before


class PeerExchange:
    """The arg-min exchange of a sharded bundle over peer-mapped memory (NVLink / NVSwitch) instead of NCCL: every rank
    owns a small mailbox all ranks of the box map through CUDA IPC.  While the group is open and the engine has a
    candidate range, ``engine.grid_launch()`` itself ends with two kernels that store the shard's record and collider
    count straight into the peers' mailboxes and merge what arrived (rp_peer_* of the C-ABI); ``engine.grid_result()``
    then returns the GLOBAL result on every rank.  ``torch.distributed`` is used once, at set-up, to hand the 64-byte
    IPC handles around."""

    def __init__(self, engine, rank=None, world=None, group=None):
        from commonroad_rp_b200._lib import RpError
        self.engine = engine
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        # every step that can fail on one rank is followed by an exchange of its outcome, so that the ranks either all
        # open the group or all raise (a rank that raised alone would leave the others waiting in a collective)
        try:
            handle, err = engine.peer_create(), None
        except RpError as exc:
            handle, err = None, str(exc)
        if self.world > 1:
            every = [None] * self.world
            dist.all_gather_object(every, handle, group=group)          # set-up only; any backend
        else:
            every = [handle]
        if any(h is None for h in every):
            engine.peer_close()
            raise RpError("peer mailbox could not be created on rank(s) %s%s"
                          % ([r for r, h in enumerate(every) if h is None], ": " + err if err else ""))
        try:
            engine.peer_open(self.rank, self.world, b"".join(every))
            err = None
        except RpError as exc:
            err = str(exc)
        if self.world > 1:
            # (doubles as the barrier: every mailbox exists and is zeroed before the first store into it)
            oks = [None] * self.world
            dist.all_gather_object(oks, err is None, group=group)
        else:
            oks = [err is None]
        if not all(oks):
            engine.peer_close()
            raise RpError("peer mailboxes could not be mapped on rank(s) %s%s"
                          % ([r for r, ok in enumerate(oks) if not ok], ": " + err if err else ""))

    def close(self):
        self.engine.peer_close()


class ScenarioBatch:
    """Independent scenarios evaluated together (BASELINE config 5: scenario-major sharding, no exchange on the data
    path).  Every scenario keeps its own device context (reference tables, obstacle tables stay resident); a
    replanning cycle of ALL scenarios is one host->device copy, four kernel launches over a global (scenario, chunk)
    work queue and one device->host copy (``_lib.Batch`` / rp_batch_* of the C-ABI)."""

    def __init__(self, device=None, stream=None):
        from commonroad_rp_b200._device import current_device_and_stream
        if device is None or stream is None:
            dev, st = current_device_and_stream()
            device = dev if device is None else device
            stream = st if stream is None else stream
        self.device, self.stream = device, stream
        self.engines = []
        self.coordinate_systems = {}
        self._batch = None
        self.x0_cart = self.x0_lon = self.x0_lat = None

    def add_scenario(self, vehicle, coordinate_system, collision_checker, proj_limit=20.0):
        """vehicle: VehicleConfiguration; coordinate_system: CoordinateSystem, or the (smoothed, de-duplicated)
        reference polyline as an (n, 2) array -- its tables are then derived on the device (rp_ctx_set_reference_polyline:
        what CoordinateSystem.__init__ computes with numpy, utils_coordinate_system.py:101-118); collision_checker:
        collision.CollisionChecker.  Returns the scenario's index."""
        import numpy as np
        from commonroad_rp_b200._lib import Engine
        eng = Engine(self.device, self.stream)
        eng.set_vehicle(vehicle.length, vehicle.width, vehicle.wb_rear_axle, vehicle.wheelbase, vehicle.a_max,
                        vehicle.v_switch, vehicle.delta_max, vehicle.v_delta_max, vehicle.kappa_max)
        if isinstance(coordinate_system, np.ndarray):
            eng.set_reference_polyline(coordinate_system, proj_limit)
        else:
            tb = coordinate_system.device_tables()
            eng.set_reference(tb["ref_pos"], tb["ref_theta"], tb["ref_curv"], tb["ref_curv_d"], tb["path_xy"],
                              tb["path_s"], tb["path_normals"], tb["proj_limit"])
        collision_checker.upload(eng)
        self.engines.append(eng)
        if self._batch is not None:
            self._batch.close()
            self._batch = None
        return len(self.engines) - 1

    def add_scenario_from(self, scenario, reference_path, vehicle, road_boundary_method="obb_rectangles", proj_limit=20.0,
                          continuous_collision_check=False):
        """A CommonRoad scenario (commonroad-io objects, or the package's own reader ``utility.scenario_io``) -> the
        scenario's device tables: static obstacles, dynamic obstacles as per-time-index boxes, the road boundary of its
        lanelet network (``ReactivePlanner.set_collision_checker``, reference reactive_planner.py:218-256), and the
        reference tables derived on the device from ``reference_path`` (smoothed and resampled on the host like
        ``CoordinateSystem.__init__``).  Returns the scenario's index."""
        from commonroad_rp_b200 import collision as rpc
        from commonroad_rp_b200.utility.utils_coordinate_system import CoordinateSystem
        cc = rpc.CollisionChecker()
        for ob in scenario.static_obstacles:
            cc.add_collision_object(rpc.create_collision_object(ob))
        for ob in scenario.dynamic_obstacles:
            tvo = rpc.create_collision_object(ob)
            if continuous_collision_check:
                tvo, err = rpc.trajectory_preprocess_obb_sum(tvo)
                if err == -1:
                    raise Exception("Invalid input for trajectory_preprocess_obb_sum: dynamic obstacle elements overlap")
            cc.add_collision_object(tvo)
        cc.add_collision_object(rpc.create_road_boundary_obstacle(scenario, method=road_boundary_method)[1])
        co = reference_path if isinstance(reference_path, CoordinateSystem) else CoordinateSystem(reference_path)
        k = self.add_scenario(vehicle, co, cc, proj_limit)
        self.coordinate_systems[k] = co
        return k

    # ---- closed loop over replanning cycles (run_planner.py:61-107 for every scenario at once) ----
    def reset(self, states_cart, low_vel_mode=0, states_curv=None):
        """``ReactivePlanner.reset(initial_state_cart=..., initial_state_curv=...)`` (reference :172-216) for all scenarios:
        Cartesian rear-axle states x0[n][6] = x, y, orientation, velocity, acceleration, steering_angle; the curvilinear
        states are taken from ``states_curv = (lon[n][3], lat[n][3])`` or projected onto each scenario's own reference
        path in ONE launch (rp_batch_initial_states; status 1 / 2 raise the reference's exceptions).  Returns (lon, lat)."""
        import numpy as np
        x0 = np.ascontiguousarray(states_cart, dtype=np.float64).reshape(len(self.engines), 6)
        if states_curv is None:
            lon, lat, status = self.batch.initial_states(x0, low_vel_mode)
            if (status == 1).any():
                raise ValueError("Initial state could not be transformed (scenarios %s)." % np.flatnonzero(status == 1).tolist())
            if (status == 2).any():
                raise Exception("Initial state or reference incorrect! Curvilinear velocity is negative (scenarios %s)"
                                % np.flatnonzero(status == 2).tolist())
        else:
            lon = np.ascontiguousarray(states_curv[0], dtype=np.float64).reshape(-1, 3)
            lat = np.ascontiguousarray(states_curv[1], dtype=np.float64).reshape(-1, 3)
        self.x0_cart, self.x0_lon, self.x0_lat = x0, lon, lat
        return lon, lat

    def advance(self, steps):
        """After ``plan``: move every scenario ``steps`` time steps along its winner (the next replanning cycle starts
        there, run_planner.py:84-107) -- one launch for all scenarios.  Updates x0_cart / x0_lon / x0_lat in place for
        the scenarios that have a winner and returns the (n, 16) state rows (column 12: 1 = advanced)."""
        w = self.batch.winner_states(steps)
        ok = w[:, 12] > 0.5
        if getattr(self, "x0_cart", None) is not None:
            self.x0_cart[ok, 0:5] = w[ok, 0:5]                     # x, y, orientation, velocity, acceleration
            self.x0_lon[ok] = w[ok, 6:9]
            self.x0_lat[ok] = w[ok, 9:12]
        return w

    def __len__(self):
        return len(self.engines)

    @property
    def batch(self):
        if self._batch is None:
            from commonroad_rp_b200._lib import Batch
            self._batch = Batch(self.engines, self.stream)
        return self._batch

    def upload(self, cycle_inputs):
        """cycle_inputs[k] = (rp_plan_inputs, t, lon, d[, traj_len]) of scenario k, or the packed form of all
        scenarios (``_lib.Batch.pack``: one call instead of one per scenario)."""
        b = self.batch
        if isinstance(cycle_inputs, dict):
            b.set_inputs_all(cycle_inputs)
            return
        for k, item in enumerate(cycle_inputs):
            b.set_inputs(k, *item)

    def launch(self):
        self.batch.launch()

    def results(self):
        return self.batch.results()

    def plan(self, cycle_inputs):
        self.upload(cycle_inputs)
        self.launch()
        return self.results()

    def plan_one_by_one(self, cycle_inputs):
        """the same cycle as one launch chain per scenario (what the batch replaces; kept for comparison)"""
        for eng, item in zip(self.engines, cycle_inputs):
            eng.grid_upload(*item)
        for eng in self.engines:
            eng.grid_launch()
        return [eng.grid_result() for eng in self.engines]

    def close(self):
        if self._batch is not None:
            self._batch.close()
            self._batch = None
        for eng in self.engines:
            eng.close()
        self.engines = []
