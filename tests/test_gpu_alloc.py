"""GPU parity of the device-side block allocator / block-state ops (kvc_alloc.cu) through the C ABI:
bit-exact against (1) the golden replay of the reference's ParallelBlockAllocator + BlockState +
BlockSpaceManagerKVC methods (tests/golden/block_manager.npz) and (2) the numpy oracle on a longer
randomized schedule at Llama-like head counts, incl. exhaustion."""
import numpy as np
import pytest
import torch

from tests.golden_util import load_cases
from tests.test_oracle_golden import assert_block_state_equal, replay_block_manager

pytestmark = pytest.mark.gpu

BLOCKMGR = load_cases("block_manager.npz")


class CudaBlockManager:
    """Adapter with the oracle's method names over capi.DeviceBlockAllocator."""

    def __init__(self, L, H, B, MS, MB, NB, via="capi"):
        from vllm_kvcompress_b200 import capi
        self.via = via
        dev = "cuda"
        self.L, self.H, self.B = L, H, B
        self.block_tables = torch.full((L, MS, H, MB), -1, dtype=torch.int32, device=dev)
        self.context_lens = torch.zeros((L, MS, H), dtype=torch.int32, device=dev)
        self.token_positions = torch.full((NB, B), -1, dtype=torch.int32, device=dev)
        self.meta = [torch.full((NB,), -1 if i == 0 else 0, dtype=torch.int32, device=dev) for i in range(4)]
        self.alloc = capi.DeviceBlockAllocator(NB, self.block_tables, self.context_lens, self.token_positions, B,
                                               metadata=tuple(self.meta))

    @staticmethod
    def _i(x):
        return torch.tensor(x, dtype=torch.int32, device="cuda")

    def add_sequences(self, slots, lens):
        if self.via == "torch":
            from vllm_kvcompress_b200 import _custom_ops as ops
            a = self.alloc
            return ops.alloc_add_sequences(a.free_mask, a.free_count, a.status, self.block_tables, self.context_lens,
                                           self._i(slots), self._i(lens), max(lens), self.token_positions, self.B,
                                           *self.meta)
        self.alloc.add_sequences(self._i(slots), self._i(lens), max(lens))

    def append_slots(self, slots, positions):
        out = torch.full((self.L, len(slots), self.H), -1, dtype=torch.int64, device="cuda")
        if self.via == "torch":
            from vllm_kvcompress_b200 import _custom_ops as ops
            a = self.alloc
            ops.alloc_append_slots(a.free_mask, a.free_count, a.status, self.block_tables, self.context_lens,
                                   self._i(slots), self._i(positions), out, self.token_positions, self.B, *self.meta)
        else:
            self.alloc.append_slots(self._i(slots), self._i(positions), out)
        return out.cpu().numpy()

    def free_compressed(self, slots, counts):
        counts = torch.as_tensor(np.ascontiguousarray(counts, dtype=np.int32)).cuda()
        if self.via == "torch":
            from vllm_kvcompress_b200 import _custom_ops as ops
            a = self.alloc
            return ops.alloc_free_compressed(a.free_mask, a.free_count, self.block_tables, self.context_lens, counts,
                                             self._i(slots), self.B, self.meta[0])
        self.alloc.free_compressed(self._i(slots), counts)

    def remove_sequences(self, slots):
        if self.via == "torch":
            from vllm_kvcompress_b200 import _custom_ops as ops
            a = self.alloc
            return ops.alloc_remove_sequences(a.free_mask, a.free_count, self.block_tables, self.context_lens,
                                              self._i(slots), self.B, self.meta[0])
        self.alloc.remove_sequences(self._i(slots))

    def state(self):
        n = lambda t: t.cpu().numpy()
        return (n(self.block_tables), n(self.context_lens), n(self.alloc.free_mask), *[n(m) for m in self.meta],
                n(self.token_positions))


@pytest.mark.parametrize("via", ["capi", "torch"])
@pytest.mark.parametrize("name", sorted(BLOCKMGR))
def test_block_manager_golden(name, via):
    c = BLOCKMGR[name]
    L, H, B, MS, MB, NB = (int(c[k]) for k in ("L", "H", "B", "MS", "MB", "NB"))
    mgr = CudaBlockManager(L, H, B, MS, MB, NB, via=via)

    def check(i, pre):
        assert int(mgr.alloc.free_count.item()) == int(c[pre + "free_count"]), f"step {i}"
        assert int(mgr.alloc.status.item()) == 0
        assert_block_state_equal(c, pre, B, *mgr.state())

    replay_block_manager(c, mgr, check)


@pytest.mark.parametrize("cfg", [(1, 32, 8, 16, 24, 40, 70000), (2, 4, 2, 16, 6, 30, 300 * 256 + 77), (3, 2, 2, 1, 3, 40, 700)])
def test_block_manager_random_vs_oracle(cfg):
    from oracle import kvc_oracle as O
    seed, L, H, B, MS, MB, NB = cfg
    rng = np.random.RandomState(seed)
    ora = O.BlockManagerOracle(L, H, B, MS, MB, NB)
    mgr = CudaBlockManager(L, H, B, MS, MB, NB)
    live, lens = [], {}
    free_slots = list(range(MS))

    def both(fn, *a):
        r0 = getattr(ora, fn)(*a)
        r1 = getattr(mgr, fn)(*a)
        return r0, r1

    for it in range(60):
        op = rng.choice(["add", "append", "append", "append", "free", "remove"])
        if op == "add" and free_slots:
            k = rng.randint(1, min(3, len(free_slots)) + 1)
            slots = [free_slots.pop(rng.randint(len(free_slots))) for _ in range(k)]
            ls = [int(rng.randint(1, (MB - 4) * B)) for _ in slots]
            both("add_sequences", slots, ls)
            for s_, n_ in zip(slots, ls):
                live.append(s_)
                lens[s_] = n_
        elif op == "append" and live:
            batch = [s_ for s_ in rng.permutation(live) if ora.context_lens[:, s_].max() < (MB - 1) * B]
            if not batch:
                continue
            for _ in range(int(rng.randint(1, 2 * B + 2))):
                if any(ora.context_lens[:, s_].max() >= MB * B - 1 for s_ in batch):
                    break
                pos = [lens[s_] for s_ in batch]
                m0, m1 = both("append_slots", batch, pos)
                np.testing.assert_array_equal(m0, m1)
                for s_ in batch:
                    lens[s_] += 1
        elif op == "free" and live:
            batch = [int(s_) for s_ in rng.permutation(live)[: rng.randint(1, len(live) + 1)]]
            nblk = (ora.context_lens[:, batch] + B - 1) // B                      # [L, batch, H]
            counts = rng.randint(0, nblk + 1).transpose(1, 0, 2).astype(np.int32)  # [batch, L, H]
            both("free_compressed", batch, counts)
        elif op == "remove" and live:
            s_ = live.pop(rng.randint(len(live)))
            free_slots.append(s_)
            both("remove_sequences", [s_])
        bt, ctx, fm, si, li, hi, lo, tp = mgr.state()
        np.testing.assert_array_equal(ctx, ora.context_lens, err_msg=f"iteration {it} ({op})")
        np.testing.assert_array_equal(fm.astype(bool), ora.free_mask)
        alloc = np.arange(MB)[None, None, None] < ((ctx + B - 1) // B)[..., None]
        np.testing.assert_array_equal(bt[alloc], ora.block_tables[alloc])
        np.testing.assert_array_equal(si, ora.seq_index_by_block)
        used = si >= 0
        np.testing.assert_array_equal(tp[used], ora.token_positions[used])
        np.testing.assert_array_equal(lo[used], ora.logical_block_num_by_block[used])
        assert int(mgr.alloc.free_count.item()) == ora.free_count
        assert int(mgr.alloc.status.item()) == 0


def test_block_manager_out_of_blocks_sets_status():
    L, H, B, MS, MB, NB = 2, 2, 16, 2, 8, 20
    mgr = CudaBlockManager(L, H, B, MS, MB, NB)
    mgr.add_sequences([0], [64])                # 2*2*4 = 16 blocks
    assert int(mgr.alloc.status.item()) == 0 and int(mgr.alloc.free_count.item()) == 4
    before = [t.copy() for t in mgr.state()]
    mgr.add_sequences([1], [48])                # needs 12 > 4: nothing may change
    assert int(mgr.alloc.status.item()) == 1
    for a, b in zip(before, mgr.state()):
        np.testing.assert_array_equal(a, b)
    mgr.alloc.status.zero_()
    mgr.append_slots([0], [64])                 # 4 heads cross a block boundary: exactly the 4 free blocks
    assert int(mgr.alloc.status.item()) == 0 and int(mgr.alloc.free_count.item()) == 0
    mgr.remove_sequences([0])
    assert int(mgr.alloc.free_count.item()) == NB
