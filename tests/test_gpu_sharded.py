"""Limb-sharded key switch (fhe_linformer_b200/sharded.py): the ranks emulated one after the other on one GPU must
reproduce the single-GPU rotation bit for bit, for every split of the extended basis."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_sharded_rotation_is_bit_exact(world):
    from fhe_linformer_b200 import Engine, sharded
    e = Engine(device=0, logN=13, L=9, dnum=3, sparse_h=64)
    sharded.register_signatures(e.lib)
    rng = np.random.default_rng(world)
    dev = torch.device("cuda", 0)
    evk = e.to_dev(np.stack([rng.integers(0, int(e.moduli[m]), e.N, dtype=np.uint64) for _ in range(e.dnum * 2) for m in range(e.L + e.K)])
                   .reshape(e.dnum, 2, e.L + e.K, e.N))
    for l in (e.L, e.L - 2, 2):
        ct = np.stack([np.stack([rng.integers(0, int(e.moduli[m]), e.N, dtype=np.uint64) for m in range(l)]) for _ in range(2)])
        g = e.galois(3)
        want = e.rotate(e.to_dev(ct), g, evk).download()
        for gather_digits in (False, True):                  # digits recomputed by every rank / transformed in shares and all-gathered
            ks = sharded.ShardedKeySwitch(e, l, sharded.LocalComm(world), device=dev, gather_digits=gather_digits)
            ks.rotate(sharded.to_tensor(ct, dev), g, evk)
            got = ks.gather_result()
            e.sync(); torch.cuda.synchronize()
            assert (got.cpu().numpy().view(np.uint64) == want).all(), (world, l, gather_digits)
            # the result stays limb-sharded: each rank's buffer holds exactly its own limbs of it
            for st in ks.states:
                mine = st.out[:, st.q_first:st.q_first + st.q_count].cpu().numpy().view(np.uint64)
                assert (mine == want[:, st.q_first:st.q_first + st.q_count]).all()
        # every limb of the extended basis belongs to exactly one rank
        owned = sorted(t for st in ks.states for first, count in st.ranges() for t in range(first, first + count))
        assert owned == list(range(l + e.K))
    e.close()
