"""Path A parity on the GPU: gc_env_step / gc_env_rollout through the C-ABI against (1) the
reference-generated golden traces and (2) the CPU oracle on seeded batches."""
import os

import numpy as np
import pytest
import torch

import gym_cooking_b200 as gcb
import oracle as O

pytestmark = pytest.mark.gpu


def _u32(t):
    return t.cpu().numpy().view(np.uint32)


from conftest import TRACE_FILES, level_source


@pytest.fixture(scope="module", params=TRACE_FILES)
def traces(golden_dir, request):
    return np.load(os.path.join(golden_dir, request.param))


def test_golden_traces_bit_exact(traces):
    """Every stored reference state (canonical form), done, reward, collision count, executed
    action and hash, for all 432 reference episodes."""
    meta, length = traces["meta"], traces["length"]
    groups = {}
    for r in range(meta.shape[0]):
        groups.setdefault((int(meta[r, 0]), int(meta[r, 1]), int(meta[r, 2])), []).append(r)
    checked = 0
    for (lvl, n_agents, max_t), rows in sorted(groups.items()):
        rows = np.array(rows)
        kb = gcb.KitchenBatch(level_source(str(traces["levels"][lvl]))[1], n_agents, len(rows), max_t, track_collisions=True)
        hash_out = torch.empty(len(rows), dtype=torch.int64, device=kb.device)
        executed = torch.empty((len(rows), n_agents), dtype=torch.uint8, device=kb.device)
        L = length[rows]
        for s in range(int(L.max()) + 1):
            if s > 0:
                acts = torch.from_numpy(traces["actions"][rows, s - 1, :n_agents].copy()).to(kb.device)
                before = kb.collisions.clone()
                kb.step(acts, hash_out=hash_out, executed_out=executed)
            live = L >= s
            st = _u32(kb.state)
            t, done, agents, keys = O.decode_batch(st, n_agents)
            assert (t[live] == traces["t"][rows, s][live]).all()
            assert (agents[live] == traces["agents"][rows, s, :n_agents][live]).all()
            assert (keys[live] == traces["keys"][rows, s][live]).all()
            assert (done[live] == traces["done"][rows, s][live]).all()
            if s > 0:
                rd = kb.reward_done.cpu().numpy()
                assert ((rd & 1)[live] == traces["done"][rows, s][live]).all()
                assert ((rd >> 1)[live] == traces["reward"][rows, s][live]).all()
                nc = (kb.collisions - before).cpu().numpy()
                assert (nc[live] == traces["ncoll"][rows, s][live]).all()
                assert (executed.cpu().numpy()[live] == traces["executed"][rows, s, :n_agents][live]).all()
                hk = hash_out.cpu().numpy().view(np.uint64)
                for r in np.nonzero(live)[0][:4]:
                    ags = [tuple(int(v) for v in a) for a in agents[r]]
                    ks = [int(k) for k in keys[r] if k != 0x3FFF]
                    assert int(hk[r]) == O.hash_canonical(int(t[r]), ags, ks)
            checked += int(live.sum())
    assert checked > (30000 if meta.shape[0] > 100 else 3000)


def test_golden_traces_plain_step(traces):
    """The same reference episodes through the PLAIN call - kb.step(actions) with no optional output,
    i.e. step2_kernel<.., EXTRAS=0, BITS=0>, the instantiation bench.py times: states (canonical form), done
    and reward at every step."""
    meta, length = traces["meta"], traces["length"]
    groups = {}
    for r in range(meta.shape[0]):
        groups.setdefault((int(meta[r, 0]), int(meta[r, 1]), int(meta[r, 2])), []).append(r)
    checked = 0
    for (lvl, n_agents, max_t), rows in sorted(groups.items()):
        rows = np.array(rows)
        kb = gcb.KitchenBatch(level_source(str(traces["levels"][lvl]))[1], n_agents, len(rows), max_t)
        assert kb.collisions is None
        L = length[rows]
        for s in range(1, int(L.max()) + 1):
            kb.step(torch.from_numpy(traces["actions"][rows, s - 1, :n_agents].copy()).to(kb.device))
            live = L >= s
            t, done, agents, keys = O.decode_batch(_u32(kb.state), n_agents)
            assert (t[live] == traces["t"][rows, s][live]).all()
            assert (agents[live] == traces["agents"][rows, s, :n_agents][live]).all()
            assert (keys[live] == traces["keys"][rows, s][live]).all()
            rd = kb.reward_done.cpu().numpy()
            assert ((rd & 1)[live] == traces["done"][rows, s][live]).all()
            assert ((rd >> 1)[live] == traces["reward"][rows, s][live]).all()
            checked += int(live.sum())
    assert checked > (30000 if meta.shape[0] > 100 else 3000)


@pytest.mark.parametrize("level,n_agents", [
    ("partial-divider_tl", 2), ("full-divider_salad", 3), ("open-divider_salad", 4), ("onion-8x8", 2), ("onion-8x8", 4),
])
def test_plain_step_matches_oracle_on_random_batches(level, n_agents):
    """cfg-2's kernel as bench.py launches it (plain step: no collision counters, hash or executed
    actions): 65536 envs x 100 uniform-random steps against the oracle, every step."""
    n = 1 << 16
    text, src = level_source(level)
    kb = gcb.KitchenBatch(src, n_agents, n, 100)
    lv = O.parse_level(text, 100)
    ost = O.reset_state(lv, n_agents, n)
    acts = kb.random_actions(100, seed=1234)
    for s in range(100):
        kb.step(acts[s])
        rd, _ = O.step_batch(lv, ost, acts[s].cpu().numpy(), n_agents, n_threads=8, want_collisions=False)
        assert (_u32(kb.state) == ost).all(), "step %d" % s
        assert (kb.reward_done.cpu().numpy() == rd).all(), "step %d" % s


def test_plain_multi_level_step_matches_oracle():
    """cfg-5's env half through the plain call: 4 agents, per-env level id over the nine levels plus the
    6-object custom level in a second batch (step2_kernel<4, 4 / 6, EXTRAS=0, MULTI=1>)."""
    n, n_agents = 9 * 2048, 4
    g = torch.Generator().manual_seed(11)
    for names in (list(gcb.levels.LEVEL_NAMES), ["onion-8x8", "open-divider_salad", "full-divider_tl"]):
        srcs = [level_source(nm) for nm in names]
        level_id = torch.randint(0, len(names), (n,), generator=g, dtype=torch.uint8)
        kb = gcb.KitchenBatch([s[1] for s in srcs], n_agents, n, 100, level_id=level_id)
        acts = kb.random_actions(100, seed=1236)
        lid = level_id.numpy()
        osts, idxs = [], []
        for l, (text, _) in enumerate(srcs):
            idx = np.nonzero(lid == l)[0]
            idxs.append(idx)
            osts.append((O.parse_level(text, 100), O.reset_state(O.parse_level(text, 100), n_agents, len(idx))))
        for s in range(100):
            kb.step(acts[s])
            a = acts[s].cpu().numpy()
            got, rd = _u32(kb.state), kb.reward_done.cpu().numpy()
            for (lv, ost), idx in zip(osts, idxs):
                ord_, _ = O.step_batch(lv, ost, a[idx], n_agents, n_threads=8, want_collisions=False)
                if s % 7 == 0 or s == 99:
                    assert (got[idx] == ost).all(), "step %d" % s
                    assert (rd[idx] == ord_).all(), "step %d" % s


@pytest.mark.parametrize("level,n_agents", [
    ("onion-8x8", 2), ("onion-8x8", 4),
    ("partial-divider_tl", 2), ("full-divider_salad", 3), ("open-divider_salad", 2), ("open-divider_salad", 4),
    ("open-divider_tomato", 1), ("full-divider_tl", 4), ("partial-divider_salad", 3),
])
def test_step_matches_oracle_on_random_batches(level, n_agents):
    """65536 envs x 100 uniform-random steps, state compared bit for bit at every step."""
    n = 1 << 16
    text, src = level_source(level)
    kb = gcb.KitchenBatch(src, n_agents, n, 100, track_collisions=True)
    lv = O.parse_level(text, 100)
    ost = O.reset_state(lv, n_agents, n)
    assert (_u32(kb.state) == ost).all()
    acts = kb.random_actions(100, seed=77)
    ocoll = np.zeros(n, dtype=np.uint32)
    for s in range(100):
        kb.step(acts[s])
        rd, coll = O.step_batch(lv, ost, acts[s].cpu().numpy(), n_agents, n_threads=8)
        ocoll += coll
        if s % 9 == 0 or s == 99:
            assert (_u32(kb.state) == ost).all(), "step %d" % s
            assert (kb.reward_done.cpu().numpy() == rd).all(), "step %d" % s
    assert (kb.collisions.cpu().numpy().view(np.uint32) == ocoll).all()
    h = kb.hash().cpu().numpy().view(np.uint64)
    assert (h[:2048] == O.hash_states(ost[:2048], n_agents)).all()


def test_rollout_equals_stepwise_and_oracle():
    """cfg-2: fused philox rollout == per-step kernel on the materialised action stream == oracle,
    including the per-step hash trace of the first 4096 envs."""
    level, n_agents, n = "partial-divider_tl", 2, 1 << 15
    a = gcb.KitchenBatch(level, n_agents, n, 100)
    b = gcb.KitchenBatch(level, n_agents, n, 100)
    trace = torch.empty((100, n), dtype=torch.int64, device=a.device)
    a.rollout(100, seed=1234, hash_trace=trace)
    acts = b.random_actions(100, seed=1234)
    hb = torch.empty(n, dtype=torch.int64, device=b.device)
    for s in range(100):
        b.step(acts[s], hash_out=hb)
        assert torch.equal(hb, trace[s]), "hash differs at step %d" % s
    assert torch.equal(a.state, b.state) and torch.equal(a.reward_done, b.reward_done)
    lv = O.parse_level(gcb.levels.level_text(level), 100)
    ost = O.reset_state(lv, n_agents, 4096)
    rd, _, oh = O.rollout_batch(lv, ost, n_agents, 100, seed=1234, n_threads=8, want_hash=True)
    assert (_u32(a.state)[:4096] == ost).all()
    assert (trace[:, :4096].cpu().numpy().view(np.uint64) == oh).all()
    # chunked rollouts (t0 offsets) and env0 offsets reproduce the same stream
    c = gcb.KitchenBatch(level, n_agents, 1024, 100)
    c.rollout(40, seed=1234, env0=2048)
    c.rollout(60, t0=40, seed=1234, env0=2048)
    assert torch.equal(c.state, a.state[2048:3072])


def test_multi_level_batch_matches_oracle():
    """cfg-5 style: 4 agents, per-env level id over all nine levels."""
    n, n_agents = 9 * 4096, 4
    g = torch.Generator().manual_seed(5)
    level_id = torch.randint(0, 9, (n,), generator=g, dtype=torch.uint8)
    kb = gcb.KitchenBatch(list(gcb.levels.LEVEL_NAMES), n_agents, n, 100, level_id=level_id)
    acts = kb.random_actions(100, seed=1236)
    for s in range(100):
        kb.step(acts[s])
    st = _u32(kb.state)
    rd = kb.reward_done.cpu().numpy()
    acts_np = acts.cpu().numpy()
    lid = level_id.numpy()
    for l, name in enumerate(gcb.levels.LEVEL_NAMES):
        idx = np.nonzero(lid == l)[0]
        lv = O.parse_level(gcb.levels.level_text(name), 100)
        ost = O.reset_state(lv, n_agents, len(idx))
        for s in range(100):
            ord_, _ = O.step_batch(lv, ost, acts_np[s][idx], n_agents, n_threads=8, want_collisions=False)
        assert (st[idx] == ost).all(), name
        assert (rd[idx] == ord_).all(), name


def test_success_episodes_stay_frozen(traces):
    """A delivered episode reports done|reward forever and its state stops changing."""
    meta = traces["meta"]
    r = next((r for r in range(meta.shape[0]) if traces["reward"][r, traces["length"][r]] == 1), None)
    if r is None:
        pytest.skip("no delivered episode in this fixture")
    lvl, n_agents, max_t = int(meta[r, 0]), int(meta[r, 1]), int(meta[r, 2])
    kb = gcb.KitchenBatch(level_source(str(traces["levels"][lvl]))[1], n_agents, 1, max_t)
    for s in range(int(traces["length"][r])):
        kb.step(torch.from_numpy(traces["actions"][r, s:s + 1, :n_agents].copy()).to(kb.device))
    assert int(kb.reward_done[0]) == 3
    snap = kb.state.clone()
    for a in range(5):
        kb.step(torch.full((1, n_agents), a, dtype=torch.uint8, device=kb.device))
        assert torch.equal(kb.state, snap) and int(kb.reward_done[0]) == 3


def test_stats_reduce():
    n = 50000
    kb = gcb.KitchenBatch("open-divider_tomato", 2, n, 30, track_collisions=True)
    kb.rollout(30, seed=3)
    s = kb.stats().cpu().numpy()
    st = _u32(kb.state)
    t = (st[:, 0] >> 24) & 127
    done = st[:, 0] >> 31
    assert s[0] == n and s[4] == int((done == 0).sum())
    assert s[1] == int(((done == 1) & (t < 30)).sum())
    assert s[2] == int(t[done == 1].sum())
    assert s[3] == int(kb.collisions.sum())
    assert (s[5:5 + 128] == np.bincount(t[done == 1], minlength=128)).all()
    # [133]: completed subtasks read off the states (random walks chop and plate now and then)
    slots = O.slots_of(st)
    mask, holder, cell = slots & 0x7F, slots >> 13, (slots >> 7) & 63
    want = 0
    for sub in kb.subtasks[0]:
        kind, _, _, goal = gcb.recipe_planner.subtask_masks(sub)
        live = holder != 7
        if kind == 3:  # Deliver: the goal object lies on the Delivery square (3, 0) -> cell 24
            hit = live & (holder == 0) & (mask == goal) & (cell == 3 * 8 + 0)
        else:
            hit = live & ((mask & goal) == goal)
        want += int(hit.any(axis=1).sum())
    assert s[133] == want and want > 0


def test_empty_and_ragged_sizes():
    for n in (1, 31, 33, 257, 1000):
        kb = gcb.KitchenBatch("full-divider_tl", 3, n, 100)
        lv = O.parse_level(gcb.levels.level_text("full-divider_tl"), 100)
        ost = O.reset_state(lv, 3, n)
        acts = kb.random_actions(20, seed=n)
        for s in range(20):
            kb.step(acts[s])
            O.step_batch(lv, ost, acts[s].cpu().numpy(), 3)
        assert (_u32(kb.state) == ost).all()
    lib = gcb._lib.load()
    lv = gcb._lib.parse_level(gcb.levels.level_text("full-divider_tl"), 100)
    import ctypes as C
    assert lib.gc_env_step(C.byref(lv), 1, None, C.c_void_p(16), C.c_void_p(16), None, None, None, None, 0, 2, None) == 0


def test_delivery_square_holding_several_objects():
    """Injected state: two dishes already lie on the Delivery square and agent-1 faces it with a third
    one (more than the reference levels can produce, but the transition is defined): every joint
    action must match the oracle."""
    level, n_agents = "partial-divider_tl", 2
    text = gcb.levels.level_text(level)
    lv = O.parse_level(text, 100)
    deliv, dish_l = 3 * 8 + 0, 2 | 8 | 32
    full = 0x7F  # every content bit: two of them overflow a 7-bit sum (the table-driven kernel adds masks)
    slots = [full | deliv << 7, full | deliv << 7, dish_l | 1 << 13, 8 | (6 * 8 + 5) << 7]  # goals stay open
    w = np.array(O.words_of((3 * 8 + 1) | (1 * 8 + 4) << 6 | 5 << 24, slots), dtype=np.uint32)
    acts = np.array([[a, b] for a in range(5) for b in range(5)], dtype=np.uint8)
    ost = np.tile(w, (25, 1))
    kb = gcb.KitchenBatch(level, n_agents, 25, 100)
    kb.state.copy_(torch.from_numpy(ost.view(np.int32)).to(kb.device))
    for rep in range(3):
        kb.step(torch.from_numpy(acts).to(kb.device))
        rd, _ = O.step_batch(lv, ost, acts, n_agents)
        assert (_u32(kb.state) == ost).all(), rep
        assert (kb.reward_done.cpu().numpy() == rd).all(), rep
    delivered = O.slots_of(ost)[:, :3]
    assert (((delivered >> 7) & 63 == deliv) & (delivered >> 13 == 0)).all(axis=1).any()  # the third dish got there


def _planes(rd, n):
    """done / reward bit planes of a reward_done byte vector, as gc_env_step_host defines them"""
    words = (n + 31) // 32
    b = np.zeros(words * 32, dtype=np.uint8)
    b[:n] = rd
    w = (1 << np.arange(32, dtype=np.uint64))
    done = ((b & 1) != 0).reshape(words, 32).astype(np.uint64) @ w
    rew = ((b & 2) != 0).reshape(words, 32).astype(np.uint64) @ w
    return np.stack([done, rew], axis=1).astype(np.uint32)


@pytest.mark.parametrize("n,kind", [(1000, "plain"), (65536 + 7, "plain"), (31, "plain"), (4099, "collisions"),
                                    (3001, "multi")])
def test_step_host_bit_planes(n, kind):
    """gc_env_step_host's bit planes (written by the step kernel itself on the plain single-level path,
    by pack_rd_kernel otherwise) equal the packed reward/done bytes, ragged tail included, and the
    states equal those of the plain device step."""
    max_t = 12  # short horizon: done bits show up within the run
    if kind == "multi":
        names = ["partial-divider_tl", "open-divider_salad"]
        lid = (torch.arange(n) % 2).to(torch.uint8)
        mk = lambda: gcb.KitchenBatch(names, 2, n, max_t, level_id=lid)
    else:
        mk = lambda: gcb.KitchenBatch("open-divider_tomato", 2, n, max_t, track_collisions=(kind == "collisions"))
    kb, ref = mk(), mk()
    acts = ref.random_actions(16, seed=21)
    words = (n + 31) // 32
    bits_dev = torch.full((words, 2), -1, dtype=torch.int32, device=kb.device)
    bits_host = torch.zeros((words, 2), dtype=torch.int32).pin_memory()
    adev = torch.empty((n, 2), dtype=torch.uint8, device=kb.device)
    seen_done = 0
    for s in range(16):
        kb.step_host(acts[s].cpu().pin_memory(), adev, None, bits_dev, bits_host)
        ref.step(acts[s])
        torch.cuda.synchronize()
        assert torch.equal(kb.state, ref.state) and torch.equal(kb.reward_done, ref.reward_done)
        expect = _planes(ref.reward_done.cpu().numpy(), n)
        assert np.array_equal(bits_host.numpy().view(np.uint32), expect), "step %d" % s
        seen_done += int(expect[:, 0].any())
    assert seen_done > 0


@pytest.mark.parametrize("level,n_agents,n", [("partial-divider_tl", 2, 70001), ("partial-divider_tl", 2, (1 << 18) + 777),
                                              ("full-divider_salad", 3, 4099),
                                              ("open-divider_salad", 4, 33333), ("onion-8x8", 1, 1000)])
def test_prepared_plans_and_joint_actions(level, n_agents, n):
    """The prepared-step path (gc_step_plan_run / _run_host) and the joint-index action format
    (j = sum_i a_i * 5^(NA-1-i); uint8, int16 for 4 agents) against gc_env_step on action bytes - which the
    tests above hold against the oracle: same states, reward/done bytes and bit planes at every step; an
    out-of-range joint index means "everybody stays".  From 2^18 envs the synchronous host call goes in two
    pipelined chunks on side streams (second parameter set)."""
    text, src = level_source(level)
    max_t = 14
    a = gcb.KitchenBatch(src, n_agents, n, max_t)                           # plan, bytes
    b = gcb.KitchenBatch(src, n_agents, n, max_t)                           # plan, joint indices
    c = gcb.KitchenBatch(src, n_agents, n, max_t)                           # plan, host joint indices -> bit planes
    ref = gcb.KitchenBatch(src, n_agents, n, max_t, track_collisions=True)  # gc_env_step (outside the plans' envelope)
    assert ref._plan(False) is None and a._plan(False) is not None
    acts = ref.random_actions(18, seed=33)
    weights = torch.tensor([5 ** (n_agents - 1 - i) for i in range(n_agents)], device=a.device)
    bits = torch.zeros(((n + 31) // 32, 2), dtype=torch.int32).pin_memory()
    for s in range(18):
        by = acts[s].clone()
        joint = (by.long() * weights).sum(1)
        if s == 3:  # out-of-range codes: bytes > 4 are "stay" per agent, a joint index >= 5^NA is "all stay"
            by[::7] = 4
            joint[::7] = 5 ** n_agents + (torch.arange(joint[::7].numel(), device=a.device) % 3)
        joint = joint.to(torch.int16 if n_agents == 4 else torch.uint8)
        ref.step(by)
        a.step(by)
        b.step(joint)
        c.step_host_bits(joint.cpu().pin_memory(), bits)
        torch.cuda.synchronize()
        for kb in (a, b, c):
            assert torch.equal(kb.state, ref.state), "step %d" % s
        assert torch.equal(a.reward_done, ref.reward_done) and torch.equal(b.reward_done, ref.reward_done)
        assert np.array_equal(bits.numpy().view(np.uint32), _planes(ref.reward_done.cpu().numpy(), n)), "step %d" % s
    assert bool(ref.done.all())  # the horizon was reached: done / timeout bits were part of the comparison


def test_table_cache_survives_more_level_sets_than_it_holds(tmp_path):
    """The device-table cache holds 64 level sets and evicts the least recently used one: 72 distinct kitchens
    (one or two extra counters on different floor squares), stepped in turn and then again from the first, all
    match the oracle - prepared plans look their tables up again at every run, so an evicted set is rebuilt."""
    base = gcb.levels.level_text("open-divider_tomato").split("\n")
    batches = []
    import itertools
    free = [sq for sq in range(20) if sq not in (11, 13)]  # (2,4) and (4,4) are the start squares of agents 4 and 3
    combos = list(itertools.combinations(free, 2))[:72]
    for k in range(72):
        rows = list(base)
        for sq in combos[k]:
            y, x = 2 + sq // 5, 1 + sq % 5    # a floor square that is not an agent start ((2,1), (4,1) are in row 1)
            rows[y] = rows[y][:x] + "-" + rows[y][x + 1:]
        text = "\n".join(rows)
        path = tmp_path / ("variant%d.txt" % k)
        path.write_text(text)
        kb = gcb.KitchenBatch(str(path), 2, 513, 100)
        lv = O.parse_level(text, 100)
        batches.append((kb, lv, O.reset_state(lv, 2, 513)))
    for rnd in range(2):
        for k, (kb, lv, ost) in enumerate(batches):
            acts = kb.random_actions(6, seed=100 * rnd + k)
            for s in range(6):
                kb.step(acts[s])
                O.step_batch(lv, ost, acts[s].cpu().numpy(), 2, want_collisions=False)
            assert (_u32(kb.state) == ost).all(), (rnd, k)
