"""Executable model (pure Python, CPU) of the synchronisation protocol of the persistent tcgen05 attention kernel
(llm-rankers_b200/csrc/attention_tc.cuh, `enc_attention_tc2_kernel`): the TMA thread, the MMA thread, the two softmax warpgroups, the
asynchronous tensor pipe and copy engine, and the twelve mbarriers between them, with the phase-parity semantics of mbarrier
(`try_wait.parity` passes once the phase with that parity has COMPLETED). The model is run under randomised schedules over random
sequences of (document, head) items (documents of 1..192 tokens: one or two query tiles) and checks what no GPU test can see directly:
  * no deadlock (every role terminates) and no barrier ever completes two phases ahead of a waiter (parity aliasing);
  * no hazard on the shared resources: Q/K and V shared-memory tiles, the P tiles, the S and O accumulators in TMEM are never
    overwritten before their last reader of the previous item has read them, and every reader sees the item it expects.
Two orders of the softmax warpgroup are modelled: MODE 0 (ships: pass 1, deferred epilogue of the previous item, pass 2) and MODE 2
(`B200RANK_ATTN=tc4`, written without GPU time: epilogue of the previous item first, then one pass over S) — the point of this file is
to vet the re-ordering of MODE 2 before it first runs on hardware. This mirrors the kernel's control flow by hand; it is a model of
the protocol, not of the arithmetic (tests/test_onepass_softmax_model.py covers that)."""
import random

import pytest


class Barrier:
    def __init__(self, name, count):
        self.name, self.count, self.pending, self.phase = name, count, count, 0   # phase = number of completed phases

    def arrive(self):
        self.pending -= 1
        assert self.pending >= 0, self.name
        if self.pending == 0:
            self.phase += 1
            self.pending = self.count


class Sim:
    def __init__(self, lens, mode, rng):
        self.lens, self.mode, self.rng = lens, mode, rng
        B = Barrier
        self.bar = dict(qk=B("bar_qk", 1), v=B("bar_v", 1), qk_free=B("qk_free", 1), v_free=B("v_free", 1))
        for t in (0, 1):
            self.bar[f"s{t}"], self.bar[f"p{t}"] = B(f"bar_s{t}", 1), B(f"bar_p{t}", 2)     # 2 = the model's two "halves" of a warpgroup
            self.bar[f"o{t}"], self.bar[f"ofree{t}"] = B(f"bar_o{t}", 1), B(f"o_free{t}", 2)
        # resources: which item they hold, and whether their last expected reader has read them
        self.qk = self.v = None
        self.P, self.S, self.O = [None, None], [None, None], [None, None]
        # per tile slot: MMA-1s / MMA-2s executed, softmax halves that drained S_t, epilogue halves that drained O_t
        self.mma1_done, self.mma2_done, self.S_drained, self.O_drained = [0, 0], [0, 0], [0, 0], [0, 0]
        self.qk_readers_left = self.v_readers_left = 0
        self.tensor_q, self.copy_q = [], []          # in-order asynchronous engines
        self.done_epilogues = []

    @staticmethod
    def ntiles(length):
        return (length + 127) >> 7

    # ---- asynchronous engines: execute queued operations in order, at random moments
    def engine_step(self, q):
        if q:
            q.pop(0)()
            return True
        return False

    # ---- roles as generators: `yield (barrier, expected completed phase index)` = wait; `yield None` = a scheduling point
    def tma(self):
        for k, length in enumerate(self.lens):
            if k > 0 and not getattr(self, "skip_qk_free_wait", False):
                yield ("qk_free", k - 1)

            def load_qk(k=k):
                assert self.qk_readers_left == 0, "Q/K overwritten before both MMA-1s of the previous item read them"
                self.qk, self.qk_readers_left = k, self.ntiles(self.lens[k])
                self.bar["qk"].arrive()
            self.copy_q.append(load_qk)
            if k > 0:
                yield ("v_free", k - 1)

            def load_v(k=k):
                assert self.v_readers_left == 0, "V overwritten before the MMA-2s of the previous item read it"
                self.v, self.v_readers_left = k, self.ntiles(self.lens[k])
                self.bar["v"].arrive()
            self.copy_q.append(load_v)
            yield None

    def mma(self):
        use = [0, 0]
        nt_prev, k = 0, 0
        n = len(self.lens)
        while k < n or nt_prev > 0:
            have = k < n
            nt_cur = self.ntiles(self.lens[k]) if have else 0
            if have:
                yield ("qk", k)
            for t in (0, 1):
                if t < nt_prev:
                    if t == 0:
                        yield ("v", k - 1)
                    yield (f"p{t}", use[t])
                    if use[t] > 0 and not getattr(self, "skip_ofree_wait", False):
                        yield (f"ofree{t}", use[t] - 1)

                    def mma2(t=t, item=k - 1, last=(t == nt_prev - 1)):
                        assert self.P[t] == item and self.v == item, f"MMA-2 operands: P {self.P[t]} V {self.v} expected {item}"
                        assert self.O_drained[t] == 2 * self.mma2_done[t], "O_t overwritten before the epilogue of its previous use read it"
                        self.mma2_done[t] += 1
                        self.v_readers_left -= 1
                        self.O[t] = item
                        self.bar[f"o{t}"].arrive()
                        if last:
                            self.bar["v_free"].arrive()
                    self.tensor_q.append(mma2)
                    use[t] += 1
                if t < nt_cur:
                    def mma1(t=t, item=k, last=(t == nt_cur - 1)):
                        assert self.qk == item, f"MMA-1 operands: Q/K hold {self.qk}, expected {item}"
                        assert self.S_drained[t] == 2 * self.mma1_done[t], "S_t overwritten before the softmax of its previous use drained it"
                        self.mma1_done[t] += 1
                        self.qk_readers_left -= 1
                        self.S[t] = item
                        self.bar[f"s{t}"].arrive()
                        if last:
                            self.bar["qk_free"].arrive()
                    self.tensor_q.append(mma1)
                yield None
            nt_prev = nt_cur
            k += 1

    def softmax(self, t, half):
        """One of the two modelled halves of warpgroup t (bar_p / o_free count 2 stands for the kernel's 128 arrivals)."""
        use, pend = 0, None
        for k, length in enumerate(self.lens):
            if t >= self.ntiles(length):
                continue

            def epilogue(item, u):
                yield (f"o{t}", u)
                assert self.O[t] == item, f"epilogue reads O_{t} of item {self.O[t]}, expected {item}"
                self.O_drained[t] += 1
                self.bar[f"ofree{t}"].arrive()
                self.done_epilogues.append((item, t, half))
            if self.mode == 2 and pend is not None:          # tc4: the previous use's epilogue comes BEFORE the wait on S_t
                yield from epilogue(*pend)
                pend = None
            yield (f"s{t}", use)
            assert self.S[t] == k, f"softmax reads S_{t} of item {self.S[t]}, expected {k}"
            yield None                                       # pass 1 (MODE 0) / the walk has started (MODE 2)
            if self.mode == 0 and pend is not None:          # ships: between the passes
                yield from epilogue(*pend)
                pend = None
            # writes into the P tile: MMA-2 of every previous use of the slot must have read it
            assert self.mma2_done[t] >= use, "P_t overwritten while MMA-2 of its previous use may still read it"
            self.P[t] = k
            yield None
            assert self.S[t] == k, "S_t changed under the softmax"
            self.S_drained[t] += 1
            self.bar[f"p{t}"].arrive()
            pend = (k, use)
            use += 1
        if pend is not None:
            yield from self.softmax_tail(t, half, pend)

    def softmax_tail(self, t, half, pend):
        item, u = pend
        yield (f"o{t}", u)
        assert self.O[t] == item
        self.O_drained[t] += 1
        self.bar[f"ofree{t}"].arrive()
        self.done_epilogues.append((item, t, half))

    def run(self):
        roles = {"tma": self.tma(), "mma": self.mma()}
        for t in (0, 1):
            for half in (0, 1):
                roles[f"wg{t}.{half}"] = self.softmax(t, half)
        waiting = {name: None for name in roles}          # pending (barrier, phase) per role
        steps = 0
        while roles:
            steps += 1
            assert steps < 200000, "livelock"
            choices = list(roles) + ["tensor", "copy"]
            self.rng.shuffle(choices)
            progressed = False
            for name in choices:
                if name == "tensor":
                    progressed = self.engine_step(self.tensor_q)
                elif name == "copy":
                    progressed = self.engine_step(self.copy_q)
                else:
                    w = waiting[name]
                    if w is not None:
                        bar, phase = w
                        if self.bar[bar].phase <= phase:
                            continue
                        assert self.bar[bar].phase <= phase + 1, f"parity aliasing: {name} waits for phase {phase} of {bar}, barrier at {self.bar[bar].phase}"
                        waiting[name] = None
                    try:
                        waiting[name] = next(roles[name])
                    except StopIteration:
                        del roles[name]
                        del waiting[name]
                    progressed = True
                if progressed:
                    break
            if not progressed:
                raise AssertionError(f"deadlock: {waiting}, tensor queue {len(self.tensor_q)}, copy queue {len(self.copy_q)}")
        while self.engine_step(self.tensor_q) or self.engine_step(self.copy_q):
            pass
        return self


@pytest.mark.parametrize("mode", [0, 2])
def test_attention_protocol_has_no_deadlock_aliasing_or_hazard(mode):
    master = random.Random(1234 + mode)
    fixed = [[184] * 12, [1], [128, 129, 128, 129, 1, 192], [192] * 5 + [64] * 5 + [192] * 5, [100] * 9, [129]]
    for trial in range(400):
        lens = fixed[trial] if trial < len(fixed) else [master.choice([1, 17, 64, 100, 128, 129, 150, 184, 192]) for _ in range(master.randint(1, 14))]
        sim = Sim(lens, mode, random.Random(master.random())).run()
        want = sorted((k, t, h) for k, L in enumerate(lens) for t in range(Sim.ntiles(L)) for h in (0, 1))
        assert sorted(sim.done_epilogues) == want, (lens, mode)
        assert sim.qk_readers_left == 0 and sim.v_readers_left == 0 and sim.O_drained == [2 * n for n in sim.mma2_done]


def test_the_model_catches_a_broken_protocol():
    """Sanity of the model itself: a TMA thread that refills Q/K without waiting for qk_free overwrites the tiles before the MMA-1s of
    the previous item have read them — the model must trip on that under some schedule. (Dropping the MMA thread's o_free wait, or the
    bar_o wait ahead of the P writes, does NOT trip it: in both orders of the softmax warpgroup those are implied by bar_p / by the
    in-order tensor pipe — the kernel keeps them as cheap belts-and-braces.)"""
    tripped = 0
    for seed in range(60):
        sim = Sim([184] * 8, 2, random.Random(seed))
        sim.skip_qk_free_wait = True
        try:
            sim.run()
        except AssertionError as e:
            assert "Q/K" in str(e), e
            tripped += 1
    assert tripped > 0
    for seed in range(60):                      # the redundant waits: removing them changes nothing the model can see
        sim = Sim([184, 100, 192, 129] * 3, 2, random.Random(seed))
        sim.skip_ofree_wait = True
        sim.run()
