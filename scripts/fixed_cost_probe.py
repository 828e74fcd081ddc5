"""Fixed cost per gc_env_step launch: device time per launch as a function of the batch size, launches
replayed from a CUDA graph (ring of independent batches, so the host never limits the rate).  With 888
resident CTAs of 256 threads, n = 227328 * k is exactly k tiles per CTA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gym_cooking_b200 as gcb

TILE = 888 * 256


def per_launch_us(n, ring=16, reps=20):  # 3 + 5 x 20 = 103 steps < horizon 127: no env is done
    kbs = [gcb.KitchenBatch("partial-divider_tl", 2, n, 127) for _ in range(ring)]
    acts = [kb.random_actions(reps + 3, seed=5 + i) for i, kb in enumerate(kbs)]
    for w in range(3):
        for r in range(ring):
            kbs[r].step(acts[r][w])
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for t in range(reps):
                for r in range(ring):
                    kbs[r].step(acts[r][3 + t])
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (4 * reps * ring)


def main():
    rows = []
    for k in (1, 2, 4, 5, 8, 32):
        n = TILE * k
        ring = 16 if k <= 8 else 4
        us = per_launch_us(n, ring=ring)
        rows.append((k, n, us))
        print("tiles/CTA %2d  n=%8d  %.2f us per launch  (%.2f us per tile)" % (k, n, us, us / k))
    for n in (1 << 20,):
        print("n=%8d (%.2f tiles/CTA)  %.2f us per launch" % (n, n / TILE, per_launch_us(n)))
    (k0, _, u0), (k1, _, u1) = rows[3], rows[-1]
    slope = (u1 - u0) / (k1 - k0)
    print("slope %.3f us per tile-iteration, intercept (fixed cost per launch) %.2f us" % (slope, u0 - slope * k0))



if __name__ == "__main__":
    main()
