"""Batched kitchen engine: N independent Overcooked episodes as one uint32[N][4] CUDA tensor.

This is the host mirror of path A (OvercookedEnvironment.reset/step,
envs/overcooked_environment.py:201-306) for a whole batch; all state mutation happens in the
kernels behind include/gymcook.h.  `envs.OvercookedEnvironment` wraps it with the reference's
gym-style single-env surface.
"""
import ctypes as C

import torch

from . import _lib, levels as _levels

# World.NAV_ACTIONS + stay (utils/world.py:16)
ACTIONS = ((0, 1), (0, -1), (-1, 0), (1, 0), (0, 0))
ACTION_INDEX = {a: i for i, a in enumerate(ACTIONS)}


class KitchenBatch:
    """N envs sharing one level (or one of several levels via `level_id`)."""

    def __init__(self, level, num_agents, num_envs, max_num_timesteps=100, device=None, level_id=None,
                 track_collisions=False):
        names = [level] if isinstance(level, str) else list(level)
        if not names:
            raise ValueError("need at least one level")
        self.level_names = names
        self.levels = [_lib.parse_level(_levels.resolve_level(nm), max_num_timesteps) for nm in names]
        self.subtasks = [self._attach_subtasks(lv) for lv in self.levels]
        self._level_arr = _lib.level_array(self.levels)
        self.n_levels = len(names)
        self.num_agents = int(num_agents)
        self.num_envs = int(num_envs)
        self.max_num_timesteps = int(max_num_timesteps)
        if not torch.cuda.is_available():
            raise _lib.GcError("gym-cooking_b200 needs a CUDA device: there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.lib = _lib.load()
        with torch.cuda.device(self.device):
            self.state = torch.empty((self.num_envs, 4), dtype=torch.int32, device=self.device)
            self.reward_done = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
            self.collisions = (torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
                               if track_collisions else None)
        if self.n_levels > 1:
            if level_id is None:
                raise ValueError("several levels need a per-env level_id tensor")
            self.level_id = level_id.to(self.device, torch.uint8).contiguous()
            # the kernels index the level tables with this byte: check it once, here
            if self.level_id.shape != (self.num_envs,):
                raise ValueError("level_id must have one entry per env (%d), got %s" % (self.num_envs, tuple(self.level_id.shape)))
            if self.num_envs and int(self.level_id.max()) >= self.n_levels:
                raise ValueError("level_id has values >= n_levels (%d)" % self.n_levels)
        else:
            self.level_id = None
        self._plans = {}  # prepared steps (gc_step_plan), by action format
        with torch.cuda.device(self.device):  # device tables now, so that a first step inside a graph capture works
            _lib.check(self.lib.gc_env_prepare(self._lv(), self.n_levels, self.num_agents))
        self.reset()

    def __del__(self):
        for plan in getattr(self, "_plans", {}).values():
            try:
                self.lib.gc_step_plan_destroy(plan)
            except Exception:
                pass

    def _plan(self, joint):
        """gc_step_plan of the plain step for this batch (one per action format), or None when the batch
        is outside the plans' envelope (several levels, collision counters)."""
        plan = self._plans.get(joint)
        if plan is None:
            if self.n_levels != 1 or self.collisions is not None or self.num_envs < 1:
                return None
            out = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self.lib.gc_step_plan_create(self._lv(), self.state.data_ptr(), self.reward_done.data_ptr(),
                                                        self.num_envs, self.num_agents,
                                                        _lib.PLAN_JOINT_ACTIONS if joint else 0, C.byref(out)))
            plan = self._plans[joint] = out.value
        return plan

    def _check_actions(self, actions, on_device):
        """-> True for the joint-index form (1-D: uint8 for <= 3 agents, int16 for 4), False for bytes"""
        joint = actions.dim() == 1
        if joint:
            want = torch.int16 if self.num_agents == 4 else torch.uint8
            ok = actions.shape[0] == self.num_envs and actions.dtype is want
        else:
            ok = actions.shape == (self.num_envs, self.num_agents) and actions.dtype is torch.uint8
        if not ok or not actions.is_contiguous() or actions.is_cuda != on_device:
            raise ValueError("actions must be a contiguous %s tensor: uint8[%d][%d], or one joint index per env "
                             "(uint8[%d], int16 for 4 agents)" % ("CUDA" if on_device else "host", self.num_envs,
                                                                  self.num_agents, self.num_envs))
        if on_device and actions.device != self.device:
            raise _lib.GcError("actions are on %s, the batch is on %s" % (actions.device, self.device))
        return joint

    @staticmethod
    def _attach_subtasks(lv):
        """Run the host recipe planner for this level and store the subtasks, in mask form, in the
        level tables (gc_level_set_subtasks) so that pairs can name them by index."""
        from . import recipe_planner as rp
        recipes = [{1: "SimpleTomato", 2: "SimpleLettuce", 3: "Salad", 4: "OnionSalad"}[lv.recipe_code[g]]
                   for g in range(lv.n_goals)]
        kinds = [{1: "Tomato", 2: "Lettuce", 4: "Onion", 8: "Plate"}[lv.object_init[k] & 0x7F]
                 for k in range(lv.n_objects)]
        subtasks = rp.level_subtasks(recipes, kinds)
        arr = (_lib.Subtask * len(subtasks))()
        for k, st in enumerate(subtasks):
            arr[k].kind, arr[k].a, arr[k].b, arr[k].goal = rp.subtask_masks(st)
        _lib.check(_lib.load().gc_level_set_subtasks(C.byref(lv), arr, len(subtasks)))
        return subtasks

    def set_subtask_masks(self, masks, level=0):
        """Replace level `level`'s subtask table by raw (kind, a, b, goal) tuples - e.g. the
        reference's own list when it kept Merge(Tomato, Lettuce) instead of Merge(Lettuce, Tomato)."""
        arr = (_lib.Subtask * len(masks))()
        for k, m in enumerate(masks):
            arr[k].kind, arr[k].a, arr[k].b, arr[k].goal = (int(v) for v in m)
        _lib.check(self.lib.gc_level_set_subtasks(C.byref(self.levels[level]), arr, len(masks)))
        self._level_arr = _lib.level_array(self.levels)

    # -- marshalling helpers ------------------------------------------------------------
    def _lv(self):
        lv = self.__dict__.get("_lv_cast")
        if lv is None or self.__dict__.get("_lv_cast_of") is not self._level_arr:
            lv = self._lv_cast = C.cast(self._level_arr, C.POINTER(_lib.Level))
            self._lv_cast_of = self._level_arr
        return lv

    def _stream(self):
        return _lib.stream_ptr(self.device)

    # -- path A ---------------------------------------------------------------------------
    def reset(self):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gc_env_reset(self._lv(), self.n_levels, _lib.ptr(self.level_id), _lib.ptr(self.state),
                                             self.num_envs, self.num_agents, self._stream()))
            self.reward_done.zero_()
            if self.collisions is not None:
                self.collisions.zero_()
        return self.state

    def step(self, actions, hash_out=None, executed_out=None):
        """actions: uint8[N][num_agents] CUDA tensor, values 0..4 - or one joint index per env
        (uint8[N], int16[N] for 4 agents: sum_i a_i * 5^(num_agents-1-i)).  In place on self.state."""
        if not actions.is_cuda:
            raise _lib.GcError("actions must be a CUDA tensor: libgymcook has no CPU path (step_host takes host actions)")
        joint = self._check_actions(actions, True)
        # hot call: at 2^20 envs the kernel takes ~9 us, so the plain step goes through a prepared plan
        # (one launch, three arguments) and the marshalling is kept to pointer reads
        dev = self.device
        if hash_out is None and executed_out is None and torch.cuda.current_device() == dev.index:
            plan = self._plan(joint)
            if plan is not None:
                rc = self.lib.gc_step_plan_run(plan, actions.data_ptr(), _lib.stream_ptr(dev))
                if rc != 0:
                    _lib.check(rc)
                return self.reward_done
        if joint:
            raise _lib.GcError("joint-index actions need the plain single-level step (no hash / executed / collision outputs)")
        guard = None
        if torch.cuda.current_device() != dev.index:
            guard = torch.cuda.device(dev)
            guard.__enter__()
        try:
            rc = self.lib.gc_env_step(
                self._lv(), self.n_levels, _lib.ptr(self.level_id), self.state.data_ptr(), actions.data_ptr(),
                self.reward_done.data_ptr(), _lib.ptr(hash_out),
                self.collisions.data_ptr() if self.collisions is not None else None, _lib.ptr(executed_out),
                self.num_envs, self.num_agents, torch.cuda.current_stream(dev).cuda_stream)
        finally:
            if guard is not None:
                guard.__exit__(None, None, None)
        if rc != 0:
            _lib.check(rc)
        return self.reward_done

    def step_host_bits(self, actions_host, bits_host, stream=None):
        """gc_step_plan_run_host: actions from a (pinned) host tensor - bytes [N][num_agents] or joint indices
        [N] - results as the done / reward bit planes in `bits_host` (pinned int32[(N+31)//32][2]); returns
        when the copies and the step are done.  None when the batch is outside the plans' envelope.
        With `stream` (a torch.cuda.Stream) the three operations are only enqueued on it
        (gc_step_plan_enqueue_host): synchronise the stream before reading `bits_host`."""
        joint = self._check_actions(actions_host, False)
        if bits_host.is_cuda or not bits_host.is_contiguous() or bits_host.numel() * bits_host.element_size() < (self.num_envs + 31) // 32 * 8:
            raise ValueError("bits_host must be a contiguous host tensor of (N+31)//32 x 2 32-bit words")
        plan = self._plan(joint)
        if plan is None:
            return None
        with torch.cuda.device(self.device):
            if stream is not None:
                _lib.check(self.lib.gc_step_plan_enqueue_host(plan, actions_host.data_ptr(), bits_host.data_ptr(),
                                                              stream.cuda_stream))
            else:
                _lib.check(self.lib.gc_step_plan_run_host(plan, actions_host.data_ptr(), bits_host.data_ptr(),
                                                          _lib.stream_ptr(self.device)))
        return bits_host

    def step_host(self, actions_host, actions_dev, reward_done_host=None, bits_dev=None, bits_host=None):
        """gc_env_step_host: actions from a (pinned) host tensor, results into pinned host memory -
        reward/done bytes (`reward_done_host`, uint8[N]) and/or bit planes (`bits_dev` / `bits_host`,
        int32[(N+31)//32][2]: done bits, reward bits) - one library call that returns when the
        copies and the step are done."""
        if actions_host.shape != (self.num_envs, self.num_agents) or actions_host.dtype is not torch.uint8 \
                or actions_host.is_cuda or not actions_host.is_contiguous():
            raise ValueError("actions_host must be a contiguous uint8 host tensor [%d, %d]" % (self.num_envs, self.num_agents))
        n_words = (self.num_envs + 31) // 32
        for name, t, numel, on_dev in (("actions_dev", actions_dev, self.num_envs * self.num_agents, True),
                                       ("reward_done_host", reward_done_host, self.num_envs, False),
                                       ("bits_dev", bits_dev, 2 * n_words, True), ("bits_host", bits_host, 2 * n_words, False)):
            if t is None:
                continue
            if t.is_cuda != on_dev or (on_dev and t.device != self.device) or not t.is_contiguous() \
                    or t.numel() * t.element_size() < numel * (1 if "bits" not in name else 4):
                raise ValueError("%s: wrong device, layout or size" % name)
        if bits_dev is not None and bits_dev.data_ptr() % 8:
            raise ValueError("bits_dev must be 8-byte aligned (the kernel stores one uint2 per 32 envs)")
        dev = self.device
        guard = None
        if torch.cuda.current_device() != dev.index:
            guard = torch.cuda.device(dev)
            guard.__enter__()
        try:
            rc = self.lib.gc_env_step_host(
                self._lv(), self.n_levels, _lib.ptr(self.level_id), self.state.data_ptr(), actions_host.data_ptr(),
                actions_dev.data_ptr(), self.reward_done.data_ptr(),
                reward_done_host.data_ptr() if reward_done_host is not None else None,
                bits_dev.data_ptr() if bits_dev is not None else None,
                bits_host.data_ptr() if bits_host is not None else None,
                self.collisions.data_ptr() if self.collisions is not None else None, self.num_envs, self.num_agents,
                torch.cuda.current_stream(dev).cuda_stream)
        finally:
            if guard is not None:
                guard.__exit__(None, None, None)
        if rc != 0:
            _lib.check(rc)
        return reward_done_host if reward_done_host is not None else bits_host

    def step_range(self, lo, hi, actions, stream=None):
        """Step only envs [lo, hi) with actions uint8[hi-lo][num_agents], on `stream` (a
        torch.cuda.Stream; default: current).  Lets a caller pipeline host copies of one chunk
        against the kernel of another."""
        if self.n_levels > 1:
            raise _lib.GcError("step_range needs a single-level batch")
        st = stream.cuda_stream if stream is not None else self._stream()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gc_env_step(
                self._lv(), 1, None, self.state.data_ptr() + 16 * lo, _lib.ptr(actions, torch.uint8),
                self.reward_done.data_ptr() + lo, None,
                (self.collisions.data_ptr() + 4 * lo) if self.collisions is not None else None, None,
                hi - lo, self.num_agents, st))

    def rollout(self, n_steps, t0=0, env0=0, seed=1234, hash_trace=None):
        """n_steps fused steps with the philox uniform-random action stream (cfg-2)."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.gc_env_rollout(
                self._lv(), self.n_levels, _lib.ptr(self.level_id), _lib.ptr(self.state), _lib.ptr(self.reward_done),
                _lib.ptr(hash_trace), _lib.ptr(self.collisions), self.num_envs, self.num_agents, int(n_steps),
                int(t0), int(env0), int(seed), self._stream()))
        return self.reward_done

    def random_actions(self, n_steps, t0=0, env0=0, seed=1234):
        """The same action stream materialised: uint8[n_steps][N][num_agents]."""
        with torch.cuda.device(self.device):
            out = torch.empty((n_steps, self.num_envs, self.num_agents), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.gc_fill_random_actions(_lib.ptr(out), self.num_envs, self.num_agents, int(n_steps),
                                                       int(t0), int(env0), int(seed), self._stream()))
        return out

    def hash(self, out=None):
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(self.num_envs, dtype=torch.int64, device=self.device)
            _lib.check(self.lib.gc_state_hash(_lib.ptr(self.state), _lib.ptr(out), self.num_envs, self.num_agents,
                                              self._stream()))
        return out

    def stats(self, out=None):
        """Episode statistics vector (int64[134], include/gymcook.h GC_STATS_LEN); `out` is
        accumulated into, so several shards / calls can share one vector."""
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.zeros(_lib.STATS_LEN, dtype=torch.int64, device=self.device)
            _lib.check(self.lib.gc_stats_reduce(_lib.ptr(self.state), _lib.ptr(self.collisions), self._lv(),
                                                self.n_levels, _lib.ptr(self.level_id), _lib.ptr(out),
                                                self.num_envs, self._stream()))
        return out

    # -- views --------------------------------------------------------------------------------
    @property
    def done(self):
        return (self.reward_done & 1).bool()

    @property
    def reward(self):
        return (self.reward_done >> 1) & 1


def slots_of_words(words):
    """The six objects of one packed state (byte planes, include/gymcook.h) as 16-bit working slots
    mask | cell << 7 | holder << 13 (holder 1..4 = held by that agent, 7 = dead: slot 0xE000)."""
    w = [int(v) & 0xFFFFFFFF for v in words]
    place = [(w[1] >> (8 * k)) & 0xFF for k in range(4)] + [(w[3] >> (8 * k)) & 0xFF for k in range(2)]
    mask = [(w[2] >> (8 * k)) & 0xFF for k in range(4)] + [(w[3] >> (8 * (k + 2))) & 0xFF for k in range(2)]
    return [(((p & 7) << 13) if p >= _lib.PLACE_HELD else (p << 7)) | m for p, m in zip(place, mask)]


def words_from_slots(w0, slots):
    """Inverse of slots_of_words: word 0 and up to six working slots -> the four packed words."""
    slots = list(slots) + [_lib.SLOT_DEAD] * (_lib.MAX_OBJECTS - len(slots))
    place = [(_lib.PLACE_HELD | (s >> 13)) if (s >> 13) else (s >> 7) & 63 for s in slots]
    mask = [s & 0x7F for s in slots]
    return [int(w0) & 0xFFFFFFFF, sum(place[k] << (8 * k) for k in range(4)), sum(mask[k] << (8 * k) for k in range(4)),
            place[4] | place[5] << 8 | mask[4] << 16 | mask[5] << 24]


def decode_state(words, num_agents):
    """One packed state (4 ints) -> dict(t, done, agents=[(x, y, hold_mask)], objects=[(mask, x, y, holder)])."""
    w = [int(v) & 0xFFFFFFFF for v in words]
    slots = slots_of_words(w)
    cells = [(w[0] >> (6 * i)) & 63 for i in range(num_agents)]
    agents = []
    for i, c in enumerate(cells):
        hold = 0
        for s in slots:
            if (s >> 13) == i + 1:
                hold = s & 0x7F
        agents.append((c & 7, c >> 3, hold))
    objects = []
    for s in slots:
        holder = s >> 13
        if holder == 7:
            continue
        c = cells[holder - 1] if holder else (s >> 7) & 63
        objects.append((s & 0x7F, c & 7, c >> 3, holder))
    return dict(t=(w[0] >> 24) & 127, done=bool(w[0] >> 31), agents=agents, objects=objects)
