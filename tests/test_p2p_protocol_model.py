"""CPU model of the peer-memory gradient exchange protocol (csrc/layers.cu: p2p_reduce_sqnorm_kernel / p2p_wait_zero_kernel; ops._P2PGrad):
one thread per rank, numpy arrays as the IPC-shared gradient buffers and flag arrays, random delays between the phases.  What it guards is the
PROTOCOL -- who waits for which flag word, the step counter that survives CUDA-graph replays, zero_grad calls without a step in between (the
warm-up of GraphedTrainer._capture) -- not the memory model (system-scope release / acquire, checked on the GPUs by scripts/check_p2p.py)."""
import random
import threading
import time

import numpy as np
import pytest

MAXW = 16


class Rank:
    def __init__(self, r, world, n, shared):
        self.r, self.world, self.n, self.sh = r, world, n, shared
        self.step_ctr = 0                     # state[0]: local
        self.gsum = np.zeros(n)

    def _wait(self, word, want, what):
        t0 = time.time()
        while self.sh["flags"][self.r][word] - want < 0:
            time.sleep(0)
            assert time.time() - t0 < 20, f"rank {self.r}: {what} never arrived"

    def zero_grad(self):                      # p2p_wait_zero_kernel
        step = self.step_ctr
        for q in range(self.world):
            self._wait(MAXW + q, step, f"done-reading({step}) of rank {q}")
        self.sh["grads"][self.r][:] = 0
        self.sh["stamp"][self.r] = -1         # buffer holds no step's gradients

    def backward(self, step, rng):            # the backward pass fills this rank's buffer
        time.sleep(rng.random() * 1e-3)
        self.sh["grads"][self.r][:] += (self.r + 1) * 1000 + step
        self.sh["stamp"][self.r] = step

    def opt_step(self, rng):                  # p2p_reduce_sqnorm_kernel
        step = self.step_ctr + 1
        for q in range(self.world):
            self.sh["flags"][q][self.r] = step                 # "my gradients of `step` are complete"
        for q in range(self.world):
            self._wait(q, step, f"ready({step}) of rank {q}")
        time.sleep(rng.random() * 1e-3)
        acc = np.zeros(self.n)
        for q in range(self.world):
            assert self.sh["stamp"][q] == step, f"rank {self.r} reads rank {q}'s buffer of step {self.sh['stamp'][q]} while reducing step {step}"
            acc += self.sh["grads"][q]
        self.gsum = acc
        for q in range(self.world):
            self.sh["flags"][q][MAXW + self.r] = step          # "I am done reading your buffer"
        self.step_ctr = step


@pytest.mark.parametrize("world", [1, 2, 5, 8])
def test_exchange_protocol_never_reads_a_stale_or_cleared_buffer(world):
    n, steps = 16, 40
    shared = {"grads": [np.zeros(n) for _ in range(world)], "flags": [np.zeros(2 * MAXW, np.int64) for _ in range(world)], "stamp": [-1] * world}
    ranks = [Rank(r, world, n, shared) for r in range(world)]
    errors, sums = [], [[] for _ in range(world)]

    def run(r):
        rng = random.Random(100 + r)
        try:
            k = ranks[r]
            for _ in range(3):                # warm-up: zero_grad without a step (GraphedTrainer._capture)
                k.zero_grad()
            for s in range(1, steps + 1):
                k.zero_grad()
                if rng.random() < 0.3:
                    time.sleep(rng.random() * 2e-3)          # a straggler
                k.backward(s, rng)
                k.opt_step(rng)
                sums[r].append(float(k.gsum[0]))
        except AssertionError as e:           # noqa
            errors.append(str(e))
    ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    [t.start() for t in ts]
    [t.join(60) for t in ts]
    assert not errors, errors[:3]
    want = [sum((q + 1) * 1000 + s for q in range(world)) for s in range(1, steps + 1)]
    for r in range(world):
        assert sums[r] == want, f"rank {r}: every rank must see the same sum of all ranks' gradients at every step"
