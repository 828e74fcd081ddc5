"""Drop-in `OvercookedEnvironment` with the reference's gym-style surface, batched underneath.

Reference: gym_cooking/envs/overcooked_environment.py - ctor(arglist) :37, reset() :201-250,
step(action_dict) :255-306 returning (obs, reward, done, info) with
info = {"t", "obs", "image_obs", "done", "termination_info"}, done() :316-363, reward() :365-376,
get_repr() :50-62, get_agent_names() :393-394, is_collision() :671-718, and the attributes
world / sim_agents / recipes / all_subtasks / agent_actions / obs_tm1 / collisions /
termination_info / successful / filename / t.

`OvercookedEnvironment(arglist)` is the single-episode drop-in for main.main_loop (main.py:85-117):
`step` takes the reference `action_dict {name: (dx, dy)}`.  `OvercookedEnvironment(arglist,
num_envs=N)` runs N episodes of the same level; `step` then takes a uint8[N][num_agents] tensor
(pinned host or CUDA) and returns reward / done tensors, `obs[i]` lazily materialises the
reference-shaped view of env i.  The facade only marshals: every transition runs in
gc_env_step; there is no CPU fallback.
"""
import copy
import os
from collections import namedtuple

import torch

from .. import _lib, engine, levels as _levels, recipe_planner
from ..utils.agent import COLORS, SimAgent
from ..utils.core import Object
from ..utils.world import World

CollisionRepr = namedtuple("CollisionRepr", "time agent_names agent_locations")

TERMINATION_TIMEOUT = "Terminating because passed {} timesteps"
TERMINATION_SUCCESS = "Terminating because all deliveries were completed"


class _ReferenceSurface:
    """Read-only helpers of the reference env that both the facade and its obs views offer
    (env:66-98 __str__, :378-390 print_agents/display/update_display, :396-473 run_recipes,
    :594-664 get_lower_bound_for_subtask_given_objs)."""

    def update_display(self):
        self.rep = self.world.update_display()
        for agent in self.sim_agents:
            x, y = agent.location
            self.rep[y][x] = str(agent)

    def __str__(self):
        self.update_display()
        return "\n".join("".join(c + " " for c in row) for row in self.rep)

    def display(self):
        print(str(self))

    def print_agents(self):
        for a in self.sim_agents:
            a.print_status()

    def run_recipes(self):
        return list(self.all_subtasks)

    def get_lower_bound_for_subtask_given_objs(self, subtask, subtask_agent_names, start_obj=None, goal_obj=None,
                                               subtask_action_obj=None):
        """env:594-664 through gc_lower_bound; the object arguments follow from the subtask
        (nav_utils.get_subtask_obj / get_subtask_action_obj) and are accepted for signature parity."""
        assert len(subtask_agent_names) <= 2, "passed in %d agents but can only do 1 or 2" % len(subtask_agent_names)
        from .. import navigation_planner
        return navigation_planner.lower_bound_pair(self, subtask, subtask_agent_names)


class EnvView(_ReferenceSurface):
    """What the reference hands out as `obs` (a copy of the env): world + sim_agents + t."""

    def __init__(self, env, words, actions=None):
        self.arglist = env.arglist
        self.t = (int(words[0]) >> 24) & 127
        st = engine.decode_state(words, env.num_agents)
        self.world = World(env.level, env.arglist)
        self.world.set_objects(st["objects"])
        self.sim_agents = []
        for i, (x, y, hold) in enumerate(st["agents"]):
            a = SimAgent("agent-%d" % (i + 1), COLORS[i], (x, y))
            if hold:
                a.holding = next(o for o in self.world.objects[Object((x, y), hold).name]
                                 if o.is_held and o.location == (x, y) and o.mask == hold)
            if actions is not None:
                a.action = engine.ACTIONS[int(actions[i])]
            self.sim_agents.append(a)
        self.recipes = env.recipes
        self.all_subtasks = env.all_subtasks
        self.agent_actions = {a.name: a.action for a in self.sim_agents}
        self._words = tuple(int(w) & 0xFFFFFFFF for w in words)
        self._env = env

    def get_repr(self):
        return self.world.get_repr() + tuple(a.get_repr() for a in self.sim_agents)

    def get_agent_names(self):
        return [a.name for a in self.sim_agents]

    def __eq__(self, other):
        return self.get_repr() == other.get_repr()

    def __copy__(self):
        return EnvView(self._env, self._words, [engine.ACTION_INDEX[a.action] for a in self.sim_agents])

    @property
    def obs_tm1(self):
        return self._env.obs_tm1

    def is_collision(self, *a, **k):
        return self._env.is_collision(*a, **k)


class Flag:
    """Lazy view of one bit of the per-env reward/done byte (GC_RD_*): nothing is computed until
    the caller asks (`.tensor()`, indexing, `bool(any())`), so a step over millions of envs does
    not pay a host-side pass it may never need."""

    def __init__(self, rd, bit, as_bool):
        self.rd, self.bit, self.as_bool = rd, bit, as_bool

    def tensor(self):
        t = (self.rd >> self.bit) & 1
        return t.bool() if self.as_bool else t

    def __getitem__(self, i):
        v = (int(self.rd[i]) >> self.bit) & 1
        return bool(v) if self.as_bool else v

    def __len__(self):
        return self.rd.shape[0]

    def any(self):
        return bool(((self.rd >> self.bit) & 1).any())

    def all(self):
        return bool(((self.rd >> self.bit) & 1).all())

    def sum(self):
        return int(((self.rd >> self.bit) & 1).sum())

    @property
    def shape(self):
        return self.rd.shape

    def __array__(self, dtype=None):
        a = self.tensor().numpy()
        return a.astype(dtype) if dtype is not None else a


class BitFlag(Flag):
    """The same lazy view over the bit-plane results of gc_env_step_host: `bits` is the pinned
    int32[(N+31)//32][2] buffer (plane 0 = done, plane 1 = reward), a quarter of the bytes of the
    per-env byte array on the way back over PCIe."""

    def __init__(self, bits, plane, n, as_bool):
        self.bits, self.plane, self.n, self.as_bool = bits, plane, n, as_bool

    def tensor(self):
        w = self.bits[:, self.plane]
        t = ((w[:, None] >> torch.arange(32, dtype=torch.int32)) & 1).reshape(-1)[:self.n]
        return t.bool() if self.as_bool else t.to(torch.uint8)

    def __getitem__(self, i):
        if i < 0:
            i += self.n
        v = (int(self.bits[i >> 5, self.plane]) >> (i & 31)) & 1
        return bool(v) if self.as_bool else v

    def __len__(self):
        return self.n

    def any(self):
        return bool((self.bits[:, self.plane] != 0).any())  # the bits past N are written as 0

    def sum(self):
        return int(self.tensor().sum())

    def all(self):
        return self.sum() == self.n

    @property
    def shape(self):
        return torch.Size([self.n])


class BatchObs:
    """Handle over the packed uint32[N][4] state; `obs[i]` builds the view of env i on demand."""

    def __init__(self, env, state):
        self._env = env
        self.state = state  # CUDA tensor (aliases the live state; clone it to keep a snapshot)

    def __len__(self):
        return self.state.shape[0]

    def __getitem__(self, i):
        return EnvView(self._env, self.state[i].tolist())


class OvercookedEnvironment(_ReferenceSurface):
    """Environment object for Overcooked (batched)."""

    metadata = {}

    def __init__(self, arglist, num_envs=1, device=None, track_collisions=None):
        self.arglist = arglist
        self.num_envs = int(num_envs)
        self.num_agents = int(arglist.num_agents)
        self.t = 0
        self.set_filename()
        self.rep = []
        self.collisions = []
        self.termination_info = ""
        self.successful = False
        self._device = device
        self._track = (self.num_envs == 1) if track_collisions is None else track_collisions
        self._kb = None
        self._pinned_rd = None
        self._dev_actions = None
        self._streams = None
        self._async_stream, self._async_pending = None, False
        self._pinned_bits = self._dev_bits = None
        # image_obs only when the reference would have a GameImage (env:240-246)
        self._atlas = None
        if getattr(arglist, "with_image_obs", False) or getattr(arglist, "record", False):
            from .. import render as _render
            gdir = os.path.join("misc", "game", "graphics")  # cwd-relative like the reference (game.py:7)
            self._atlas = _render.load_atlas(gdir) if os.path.isdir(gdir) else _render.default_atlas()

    def set_filename(self):  # env:116-128
        a = self.arglist
        self.filename = "{}_agents{}_seed{}".format(a.level, a.num_agents, getattr(a, "seed", 1))
        for k in (1, 2, 3, 4):
            m = getattr(a, "model%d" % k, None)
            if m is not None:
                self.filename += "_model{}-{}".format(k, m)

    # -- reset ---------------------------------------------------------------------------
    def reset(self):
        a = self.arglist
        max_t = int(getattr(a, "max_num_timesteps", 100) or 0)
        if self._kb is None:
            self._kb = engine.KitchenBatch(a.level, self.num_agents, self.num_envs, max_t, device=self._device,
                                           track_collisions=self._track)
            self.level = self._kb.levels[0]
            text = _levels.resolve_level(a.level)
            blocks = text.split("\n\n")
            self.recipes = [ln for ln in blocks[1].split("\n") if ln]
            kinds = []
            for k in range(self.level.n_objects):
                m = self.level.object_init[k] & 0x7F
                kinds.append({1: "Tomato", 2: "Lettuce", 4: "Onion", 8: "Plate"}[m])
            self.all_subtasks = recipe_planner.level_subtasks(self.recipes, kinds,
                                                              int(getattr(a, "max_num_subtasks", 14)))
            self._executed = torch.empty((self.num_envs, self.num_agents), dtype=torch.uint8, device=self._kb.device)
            self._pinned_rd = torch.empty(self.num_envs, dtype=torch.uint8).pin_memory()
            self._dev_actions = torch.empty((self.num_envs, self.num_agents), dtype=torch.uint8, device=self._kb.device)
        else:
            self._kb.reset()
        self.t = 0
        self.collisions = []
        self.termination_info = ""
        self.successful = False
        self.agent_actions = {}
        self._sync_view()
        self.obs_tm1 = copy.copy(self._view) if self.num_envs == 1 else None
        self.game = None
        if self._atlas is not None and self.num_envs == 1:  # env:240-248
            from ..misc.game.gameimage import GameImage
            self.game = GameImage(self.filename, self.get_image_obs, record=bool(getattr(a, "record", False)))
            if self.game.record:
                self.game.save_image_obs(self.t)
        return self._obs()

    def close(self):
        return

    # -- step ----------------------------------------------------------------------------
    def step(self, action_dict):
        """Single env: `action_dict {agent name: (dx, dy)}` -> (obs, reward, done, info) exactly as
        env:255-306.  Batched: uint8[N][num_agents] tensor -> (BatchObs, reward[N], done[N], info)."""
        kb = self._kb
        if isinstance(action_dict, dict):
            if self.num_envs != 1:
                raise ValueError("an action_dict drives a single env; pass a [N, num_agents] tensor")
            names = self.get_agent_names()
            acts = torch.tensor([[engine.ACTION_INDEX[tuple(action_dict[nm])] for nm in names]], dtype=torch.uint8)
        else:
            acts = action_dict
        self.t += 1
        if self.num_envs > 1:
            return self._step_batched(acts)
        if acts.is_cuda:
            dev_acts = acts
        else:
            self._dev_actions.copy_(acts, non_blocking=True)  # H2D (async when `acts` is pinned)
            dev_acts = self._dev_actions
        if self.num_envs == 1:
            words_before = kb.state[0].tolist()
            coll_before = int(kb.collisions[0]) if kb.collisions is not None else 0
        kb.step(dev_acts, executed_out=self._executed)
        self._pinned_rd.copy_(kb.reward_done, non_blocking=True)  # D2H of the step's result
        torch.cuda.current_stream(kb.device).synchronize()
        rd = self._pinned_rd
        # single env: rebuild the reference-shaped bookkeeping
        executed = self._executed[0].tolist()
        names = self.get_agent_names()
        self.obs_tm1 = EnvView(self, words_before, executed)  # state before, actions after collisions (env:273)
        if kb.collisions is not None and int(kb.collisions[0]) > coll_before:
            self._record_collisions(words_before, acts[0].tolist())
        self._sync_view()
        for i, nm in enumerate(names):
            self.agent_actions[nm] = engine.ACTIONS[executed[i]]
            self._view.sim_agents[i].action = engine.ACTIONS[executed[i]]
        if self.game is not None and self.game.record:  # env:285-286
            self.game.save_image_obs(self.t)
        done = bool(int(rd[0]) & 1)
        self.successful = bool(int(rd[0]) & 2)
        self._set_termination(done)
        new_obs = self._obs()
        info = {"t": self.t, "obs": new_obs, "image_obs": self.get_image_obs(), "done": done,
                "termination_info": self.termination_info}
        return new_obs, self.reward(), done, info

    # chunks of a large batch are pipelined over this many streams: the host->device copy of
    # chunk k+1, the kernel of chunk k and the device->host copy of chunk k-1 overlap (PCIe is
    # full duplex and the copy engines run beside the SMs)
    # (opt-in: on the measured hosts the extra launches and syncs cost more than the overlap wins,
    # 202 us vs 99 us per 2^20-env step - scripts/e2e_probe.py - so the default is one stream)
    # results of a host-actions step come back as two bit planes (done, reward) instead of one byte
    # per env: 256 KB instead of 1 MB per 2^20-env step over PCIe; `done` / `reward` unpack lazily
    PACKED_RESULTS = os.environ.get("GC_E2E_BYTE_RESULTS") is None
    PIPELINE_CHUNKS = 1
    PIPELINE_MIN_ENVS = 1 << 16

    # -- vector-env style asynchronous step (gym.vector.VectorEnv.step_async / step_wait) ---------------
    def step_async(self, actions):
        """Enqueue one step of a batched env with HOST actions (pinned uint8[N][n_agents], or joint indices
        uint8[N] / int16[N]) on this env's own stream and return at once; `step_wait()` delivers the result.
        Two envs stepped alternately overlap one's PCIe copies with the other's kernel."""
        kb = self._kb
        if self.num_envs == 1 or actions.is_cuda:
            raise ValueError("step_async is the batched, host-actions form; use step()")
        if self._async_stream is None:
            self._async_stream = torch.cuda.Stream(device=kb.device)
            self._async_stream.wait_stream(torch.cuda.current_stream(kb.device))
        if self._pinned_bits is None:
            words = (self.num_envs + 31) // 32
            self._pinned_bits = torch.zeros((words, 2), dtype=torch.int32).pin_memory()
            self._dev_bits = torch.zeros((words, 2), dtype=torch.int32, device=kb.device)
        if kb.step_host_bits(actions, self._pinned_bits, stream=self._async_stream) is None:
            raise _lib.GcError("step_async needs a single-level batch without collision counters")
        self.t += 1
        self._async_pending = True

    def step_wait(self):
        """-> (obs, reward, done, info) of the step enqueued by step_async"""
        if not self._async_pending:
            raise RuntimeError("step_wait without step_async")
        self._async_stream.synchronize()
        self._async_pending = False
        n = self.num_envs
        obs = self._obs()
        done, reward = BitFlag(self._pinned_bits, 0, n, True), BitFlag(self._pinned_bits, 1, n, False)
        return obs, reward, done, {"t": self.t, "obs": obs, "image_obs": None, "done": done, "termination_info": ""}

    def _step_batched(self, acts):
        kb = self._kb
        n = self.num_envs
        rd = self._pinned_rd
        if acts.is_cuda:
            kb.step(acts)
            rd.copy_(kb.reward_done, non_blocking=True)
            torch.cuda.current_stream(kb.device).synchronize()
        elif self.PIPELINE_CHUNKS <= 1 or n < self.PIPELINE_MIN_ENVS or kb.n_levels > 1:
            # host actions in, reward/done bytes out: one library call (gc_env_step_host) that copies
            # in (async when `acts` is pinned), steps, copies out and waits for the stream
            joint = acts.dim() == 1  # one joint index per env (engine.KitchenBatch.step): fewer bytes over PCIe
            if acts.is_contiguous() and (joint or (acts.dtype is torch.uint8 and acts.shape == (n, self.num_agents))):
                if self.PACKED_RESULTS or joint:
                    if self._pinned_bits is None:
                        words = (n + 31) // 32
                        self._pinned_bits = torch.zeros((words, 2), dtype=torch.int32).pin_memory()
                        self._dev_bits = torch.zeros((words, 2), dtype=torch.int32, device=kb.device)
                    if kb.step_host_bits(acts, self._pinned_bits) is None:  # outside the prepared-step envelope
                        kb.step_host(acts, self._dev_actions, None, self._dev_bits, self._pinned_bits)
                    obs = self._obs()
                    done = BitFlag(self._pinned_bits, 0, n, True)
                    reward = BitFlag(self._pinned_bits, 1, n, False)
                    return obs, reward, done, {"t": self.t, "obs": obs, "image_obs": None, "done": done,
                                               "termination_info": ""}
                kb.step_host(acts, self._dev_actions, rd)
            else:
                self._dev_actions.copy_(acts, non_blocking=True)
                kb.step(self._dev_actions)
                rd.copy_(kb.reward_done, non_blocking=True)
                torch.cuda.current_stream(kb.device).synchronize()
        else:
            if self._streams is None:
                self._streams = [torch.cuda.Stream(device=kb.device) for _ in range(self.PIPELINE_CHUNKS)]
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(kb.device))
            c = self.PIPELINE_CHUNKS
            for k, st in enumerate(self._streams):
                lo, hi = n * k // c, n * (k + 1) // c
                st.wait_event(ready)
                with torch.cuda.stream(st):
                    self._dev_actions[lo:hi].copy_(acts[lo:hi], non_blocking=True)
                    kb.step_range(lo, hi, self._dev_actions[lo:hi], stream=st)
                    rd[lo:hi].copy_(kb.reward_done[lo:hi], non_blocking=True)
            for st in self._streams:
                st.synchronize()
        done, reward = Flag(rd, 0, True), Flag(rd, 1, False)
        obs = self._obs()
        info = {"t": self.t, "obs": obs, "image_obs": None, "done": done, "termination_info": ""}
        return obs, reward, done, info

    def get_image_obs(self, envs=None):
        """GameImage.get_image_obs (misc/game/gameimage.py:31-51): uint8[H*80][W*80][3] of the single
        env, or uint8[m][...] of `envs` for a batch; None when image observations are off."""
        if self._atlas is None:
            return None
        from .. import render as _render
        if not isinstance(self._atlas, torch.Tensor):
            self._atlas = torch.from_numpy(self._atlas).to(self._kb.device)
        img = _render.render(self._kb, self._atlas, envs)
        return img[0].cpu().numpy() if (self.num_envs == 1 and envs is None) else img

    def _set_termination(self, done):
        max_t = int(getattr(self.arglist, "max_num_timesteps", 100) or 0)
        if done and not self.successful:
            self.termination_info = TERMINATION_TIMEOUT.format(max_t)
        elif done:
            self.termination_info = TERMINATION_SUCCESS
        else:
            self.termination_info = ""

    def _record_collisions(self, words_before, actions):
        """CollisionRepr list of env:747-752, recomputed on the host for the single-env facade
        from the pre-step state (the kernel only counts them)."""
        st = engine.decode_state(words_before, self.num_agents)
        names = self.get_agent_names()
        locs = [(x, y) for (x, y, _) in st["agents"]]
        for i in range(self.num_agents):
            for j in range(i + 1, self.num_agents):
                ex = self.is_collision(locs[i], locs[j], engine.ACTIONS[actions[i]], engine.ACTIONS[actions[j]])
                if not all(ex):
                    self.collisions.append(CollisionRepr(time=self.t, agent_names=[names[i], names[j]],
                                                         agent_locations=[locs[i], locs[j]]))

    # -- queries ---------------------------------------------------------------------------
    def done(self):
        rd = self._kb.reward_done
        if self.num_envs == 1:
            d = bool(int(rd[0]) & 1)
            self.successful = bool(int(rd[0]) & 2)
            self._set_termination(d)
            return d
        return (rd & 1).bool()

    def reward(self):
        if self.num_envs == 1:
            return 1 if self.successful else 0
        return (self._kb.reward_done >> 1) & 1

    def get_repr(self):
        return self._view.get_repr()

    def get_agent_names(self):
        return ["agent-%d" % (i + 1) for i in range(self.num_agents)]

    def is_collision(self, agent1_loc, agent2_loc, agent1_action, agent2_action):
        """env:671-718 on the static map (host helper for planners and the collision log)."""
        def nxt(loc, act):
            n = (loc[0] + act[0], loc[1] + act[1])
            return loc if self.world.loc_to_gridsquare[n].collidable else n
        n1, n2 = nxt(agent1_loc, agent1_action), nxt(agent2_loc, agent2_action)
        execute = [True, True]
        if n1 == n2:
            if n1 == agent1_loc and tuple(agent1_action) != (0, 0):
                execute[1] = False
            elif n2 == agent2_loc and tuple(agent2_action) != (0, 0):
                execute[0] = False
            else:
                execute = [False, False]
        elif agent1_loc == n2 and agent2_loc == n1:
            execute = [False, False]
        return execute

    # -- internals -------------------------------------------------------------------------
    def _sync_view(self):
        self._view = EnvView(self, self._kb.state[0].tolist())
        self.world = self._view.world
        self.sim_agents = self._view.sim_agents

    def _obs(self):
        if self.num_envs == 1:
            return copy.copy(self._view)
        return BatchObs(self, self._kb.state)

    @property
    def state(self):
        """packed uint32[N][4] CUDA tensor (include/gymcook.h)."""
        return self._kb.state

    @property
    def batch(self):
        return self._kb
