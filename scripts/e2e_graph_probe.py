"""e2e step (pinned actions in, reward/done bytes out): stream calls vs CUDA-graph replay (scratch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb

n, na = 1 << 20, 2
kb = gcb.KitchenBatch("partial-divider_tl", na, n, 100)
acts_dev = kb.random_actions(8, seed=3)
host_acts = [a.cpu().pin_memory() for a in acts_dev]
dev_a = torch.empty((n, na), dtype=torch.uint8, device=kb.device)
pin_a = torch.empty((n, na), dtype=torch.uint8).pin_memory()
pin_rd = torch.empty(n, dtype=torch.uint8).pin_memory()
st = torch.cuda.current_stream()

def plain(a):
    dev_a.copy_(a, non_blocking=True)
    kb.step(dev_a)
    pin_rd.copy_(kb.reward_done, non_blocking=True)
    st.synchronize()

def bench(fn, iters=300):
    for i in range(20): fn(host_acts[i % 8])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(iters): fn(host_acts[i % 8])
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / iters * 1e6

print("plain stream calls: %.1f us/step" % bench(plain))
# graph: fixed pinned input buffer -> device -> kernel -> pinned output
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(side):
    dev_a.copy_(pin_a, non_blocking=True); kb.step(dev_a); pin_rd.copy_(kb.reward_done, non_blocking=True)
    side.synchronize()
    with torch.cuda.graph(g, stream=side):
        dev_a.copy_(pin_a, non_blocking=True)
        kb.step(dev_a)
        pin_rd.copy_(kb.reward_done, non_blocking=True)
def graphed(a):
    pin_a.copy_(a)          # host memcpy 2 MB into the graph's input buffer (a real caller writes there directly)
    g.replay()
    torch.cuda.synchronize()
def graphed_inplace(a):
    g.replay()
    torch.cuda.synchronize()
print("graph replay (+2 MB host memcpy): %.1f us/step" % bench(graphed))
print("graph replay (caller writes the pinned buffer itself): %.1f us/step" % bench(graphed_inplace))
# pieces
def h2d(a): dev_a.copy_(a, non_blocking=True); st.synchronize()
def d2h(a): pin_rd.copy_(kb.reward_done, non_blocking=True); st.synchronize()
def kern(a): kb.step(dev_a); st.synchronize()
print("H2D 2 MB alone: %.1f us, D2H 1 MB alone: %.1f us, kernel alone: %.1f us" % (bench(h2d), bench(d2h), bench(kern)))
# zero-copy: the kernel reads the pinned host actions / writes the pinned host reward_done itself
import ctypes as C
from gym_cooking_b200 import _lib
lib = _lib.load()
def zero_copy(a, rd_host=True):
    _lib.check(lib.gc_env_step(kb._lv(), 1, None, kb.state.data_ptr(), a.data_ptr(),
                               pin_rd.data_ptr() if rd_host else kb.reward_done.data_ptr(), None, None, None, n, na, kb._stream()))
    st.synchronize()
ref = gcb.KitchenBatch("partial-divider_tl", na, n, 100)
kb.reset()
for i in range(8):
    zero_copy(host_acts[i]); ref.step(acts_dev[i])
torch.cuda.synchronize()
print("zero-copy result equal:", bool(torch.equal(kb.state, ref.state)), bool(torch.equal(pin_rd, ref.reward_done.cpu())))
print("zero-copy actions in + reward/done out: %.1f us/step" % bench(zero_copy))
def zc_in_only(a):
    zero_copy(a, rd_host=False)
    pin_rd.copy_(kb.reward_done, non_blocking=True); st.synchronize()
print("zero-copy actions in, D2H copy out: %.1f us/step" % bench(zc_in_only))
