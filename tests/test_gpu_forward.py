"""GPU parity of the encrypted Linformer forward (FHEController + host/linformer.cpp on the CUDA engine) against the slot
simulator: every checkpoint slot by slot, the 20 logits within the north star's 1e-3 absolute tolerance, identical class.
Reference parameters (N = 2^15, 28 limbs, dnum 4, 2^14 slots), S = 129 rows (the circuit's minimum)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3          # BASELINE.json north_star: decrypted logits within 1e-3 absolute, identical predicted class
CHECKPOINT_TOL = 2e-4     # intermediate ciphertexts; the loosest is tanh (slope 50) right after a bootstrap


@pytest.fixture(scope="module")
def setup(tmp_path_factory):
    from fhe_linformer_b200 import host, synth
    root = str(tmp_path_factory.mktemp("linformer"))
    model = synth.make_model(n_classes=8)
    sample = synth.make_sample(model, 128, seed=20261018 + 1)
    dirs = synth.write_files(root, model, sample)
    fc = host.FHEController(root=root).generate()
    yield fc, model, sample, dirs, root
    fc.close()


def test_forward_matches_slot_simulator(setup):
    from oracle import linformer_sim as ls
    fc, model, sample, dirs, _ = setup
    fc.ckks.ledger(True); fc.ckks.ledger_reset()
    got = {}
    logits, stages, S = fc.forward(dirs, dead_work=True, checkpoints=got)
    ref_cp = {}
    ref = ls.sim_forward(model, sample, ref_cp)
    assert S == 129
    assert len(got) >= 20
    for name, (slots, level) in got.items():
        assert np.abs(slots - ref_cp[name]).max() < CHECKPOINT_TOL, name
    assert np.abs(logits - ref).max() < LOGIT_TOL
    assert int(np.argmax(logits)) == int(np.argmax(ref))
    # the op ledger sees the reference's operation census (13 637 rotations at S = 129, plus the
    # rotations inside the 8 bootstraps)
    led = fc.ckks.ledger_dump()
    rotations = sum(n for k, (n, _) in led.items() if k.startswith("rotate@"))
    assert 13637 <= rotations < 13637 + 8 * 120
    fc.ckks.ledger(False)
    setup[0].__dict__["_faithful_logits"] = logits


def test_lean_mode_gives_the_same_logits(setup):
    fc, model, sample, dirs, _ = setup
    lean, _, _ = fc.forward(dirs, dead_work=False)
    full = fc.__dict__.get("_faithful_logits")
    if full is None:
        full, _, _ = fc.forward(dirs, dead_work=True)
    # same circuit minus dead operations; keys and encryption noise are fresh OS randomness in each run (measured 6e-5 .. 1.3e-4),
    # well inside the 1e-3 logits bar of BASELINE.json
    assert np.abs(lean - full).max() < 4e-4


def test_packed_mode_matches_slot_simulator(setup):
    """Packed mode (BASELINE north star: BSGS diagonal ct x pt matmul behind the FFN linears): 128 rows per ciphertext from the
    first affine to the second, FHEController::packed_linear for every 128 x 128 weight block.  Same checkpoints, same logits
    (<= 1e-3, same class) as the slot simulator of the reference circuit, with an order of magnitude fewer rotations."""
    from oracle import linformer_sim as ls
    fc, model, sample, dirs, _ = setup
    fc.set_option("packed_keys", 1)
    fc.ckks.ledger(True); fc.ckks.ledger_reset()
    got = {}
    logits, stages, S = fc.forward(dirs, packed=True, checkpoints=got)
    led = fc.ckks.ledger_dump(); fc.ckks.ledger(False)
    ref_cp = {}
    ref = ls.sim_forward(model, sample, ref_cp)
    for name in ("packed_hidden_block0", "packed_gelu_block0", "packed_ffn_0", "affine2_0", "encoder_out"):
        assert name in got, name
    for name, (slots, level) in got.items():
        assert np.abs(slots - ref_cp[name]).max() < CHECKPOINT_TOL, name
    assert np.abs(logits - ref).max() < LOGIT_TOL and int(np.argmax(logits)) == int(np.argmax(ref))
    rotations = sum(n for k, (n, _) in led.items() if k.startswith("rotate@"))
    assert rotations < 2500, rotations          # 13 637 + bootstraps in the faithful circuit at S = 129
    # packed + lean: only what the logits read -- the half with the CLS row through the FFN, the second affine and the column mask
    # folded into the W2 diagonals, no refresh between GELU and the pooler (whose own bootstrap then runs on the last limb)
    fc.ckks.ledger(True); fc.ckks.ledger_reset()
    got_lean = {}
    lean, _, _ = fc.forward(dirs, packed=True, dead_work=False, checkpoints=got_lean)
    led_lean = fc.ckks.ledger_dump(); fc.ckks.ledger(False)
    assert "packed_affine2_cls" in got_lean and "affine2_0" not in got_lean
    for name, (slots, level) in got_lean.items():
        assert np.abs(slots - ref_cp[name]).max() < CHECKPOINT_TOL, name
    assert np.abs(lean - ref).max() < LOGIT_TOL and int(np.argmax(lean)) == int(np.argmax(ref))
    rotations_lean = sum(n for k, (n, _) in led_lean.items() if k.startswith("rotate@"))
    assert rotations_lean < rotations, (rotations_lean, rotations)      # one bootstrap and the second half's transforms fewer
    # a packed forward needs the keys of its transforms: a controller without them fails loudly instead of generating keys on the fly
    from fhe_linformer_b200 import host
    bare = host.FHEController(root=setup[4]).generate()
    with pytest.raises(RuntimeError, match="no evaluation key"):
        bare.forward(dirs, packed=True)
    bare.close()


def test_several_samples_per_call(setup, tmp_path):
    """flh_forward_many (BASELINE config 5, ciphertext-parallel over samples): three different samples through ONE packed forward,
    every ciphertext carrying one element per sample -- each sample's logits equal its own slot simulation (<= 1e-3, same class)
    and the single-sample packed run."""
    from fhe_linformer_b200 import synth
    from oracle import linformer_sim as ls
    fc, model, sample, dirs, root = setup
    fc.set_option("packed_keys", 1)
    samples = [sample] + [synth.make_sample(model, 128, seed=777 + i) for i in range(2)]
    dirs_list = [dirs]
    for i, sm in enumerate(samples[1:]):
        d = {"weights": dirs["weights"], "input": str(tmp_path / ("input%d" % i)), "tokens": str(tmp_path / ("tokens%d" % i))}
        synth.write_sample_files(d["input"], d["tokens"], sm)
        dirs_list.append(d)
    for lean in (False, True):
        got, S = fc.forward_many(dirs_list, dead_work=not lean, packed=True)
        assert S == 129 and got.shape == (3, 20)
        for sm, z in zip(samples, got):
            ref = ls.sim_forward(model, sm)
            assert np.abs(z - ref).max() < LOGIT_TOL and int(np.argmax(z)) == int(np.argmax(ref))
    single, _, _ = fc.forward(dirs_list[1], packed=True, dead_work=False)
    assert np.abs(single - got[1]).max() < 4e-4
    with pytest.raises(RuntimeError, match="packed mode only"):
        fc.forward_many(dirs_list, packed=False)


def test_encrypted_projection_variant(setup):
    """X_E / X_F computed on the server from the encrypted rows (fl_linear_wsum) instead of uploaded by the client."""
    from oracle import linformer_sim as ls
    fc, model, sample, dirs, _ = setup
    got = {}
    logits, stages, S = fc.forward(dirs, dead_work=False, checkpoints=got, encrypted_projection=True)
    ref_cp = {}
    ref = ls.sim_forward(model, sample, ref_cp)
    assert "Projection" in stages and "projected_E0" in got
    assert np.abs(got["projected_E0"][0] - ref_cp["projected_E0"]).max() < 1e-6
    assert np.abs(ref_cp["projected_E0"] - ls.SlotSim().expanded(sample["XE"][0])).max() < 1e-12   # equals the client-side projection
    for name, (slots, level) in got.items():
        assert np.abs(slots - ref_cp[name]).max() < CHECKPOINT_TOL, name
    assert np.abs(logits - ref).max() < LOGIT_TOL and int(np.argmax(logits)) == int(np.argmax(ref))


def test_all_token_attention_variant(setup):
    """the circuit of the reference's main_2.cpp (every row attends; matmulScores(vector), 1/x on [-1, 190000])."""
    from oracle import linformer_sim as ls
    fc, model, sample, dirs, _ = setup
    got = {}
    logits, stages, S = fc.forward(dirs, dead_work=True, checkpoints=got, all_tokens=True)
    ref_cp = {}
    ref = ls.sim_forward(model, sample, ref_cp, all_tokens=True)
    assert "all_scores_exp_0" in got and "attention_row1" in got
    for name, (slots, level) in got.items():
        assert np.abs(slots - ref_cp[name]).max() < CHECKPOINT_TOL, name
    assert np.abs(logits - ref).max() < LOGIT_TOL and int(np.argmax(logits)) == int(np.argmax(ref))


def test_key_files_roundtrip_and_resume(setup):
    """generate_context(serialize) / load_context / rotation-key file / ciphertext checkpoint (F.cpp:59-89,184-301,1360-1394)."""
    from fhe_linformer_b200 import host
    fc, model, sample, dirs, root = setup
    c = fc.ckks
    v = np.random.default_rng(5).uniform(-1, 1, 16384)
    ct = c.encrypt(v)
    c.save(ct, root + "/checkpoint/x.bin")
    c.save_keys(root + "/keys/all.bin")
    other = host.FHEController(root=root)
    other.hl.flh_generate(other.h, 0, None, 0, 16384, 0)   # fresh context, different key pair would be generated from the same seed
    other.ckks = host._Borrowed(other.hl.flh_native(other.h))
    other.ckks.load_keys(root + "/keys/all.bin")
    back = other.ckks.load(root + "/checkpoint/x.bin")
    assert np.abs(other.ckks.decrypt(back) - v).max() < 1e-8
    r = other.ckks.rotate(back, 4)
    assert np.abs(other.ckks.decrypt(r) - np.roll(v, -4)).max() < 1e-8
    other.close()


def test_maximum_rows_forward_S256(tmp_path):
    """The circuit's upper bound, S = 256 rows (two full halves of 128): logits against the slot simulator, and the row-count
    guards at both ends (the reference simply misbehaves outside 129..256)."""
    from fhe_linformer_b200 import host, synth
    from oracle import linformer_sim as ls
    model = synth.make_model(n_classes=20)
    sample = synth.make_sample(model, 255, seed=99)
    dirs = synth.write_files(str(tmp_path), model, sample)
    fc = host.FHEController(root=str(tmp_path)).generate()
    logits, stages, S = fc.forward(dirs, dead_work=False)
    ref = ls.sim_forward(model, sample)
    assert S == 256
    assert np.abs(logits - ref).max() < LOGIT_TOL and int(np.argmax(logits)) == int(np.argmax(ref))
    with pytest.raises(RuntimeError, match="129 <= S <= 256"):
        fc.forward(dirs, token_limit=100)           # S = 101: below the two-half packing
    fc.close()


def test_two_controllers_in_worker_threads(setup, tmp_path_factory):
    """Throughput mode (bench.py forward.throughput_mode): a second controller with its own engine / stream / keys, both driven from
    worker threads at the same time (every C-ABI entry binds the calling thread to its engine's device).  The forwards must not
    disturb each other: each gives the logits it gives alone (same class, fresh encryption noise only)."""
    import threading
    from fhe_linformer_b200 import host
    fc, model, sample, dirs, _ = setup
    alone = fc.__dict__.get("_faithful_logits")
    if alone is None:
        alone, _, _ = fc.forward(dirs, dead_work=True)
    other = host.FHEController(root=str(tmp_path_factory.mktemp("linformer2"))).generate()
    try:
        out = {}
        def work(name, c):
            out[name] = c.forward(dirs, dead_work=True)[0]
        th = [threading.Thread(target=work, args=("a", fc)), threading.Thread(target=work, args=("b", other))]
        for t in th: t.start()
        for t in th: t.join()
        assert set(out) == {"a", "b"}
        for name in out:
            assert np.abs(out[name] - alone).max() < 4e-4, name   # fresh keys / noise per controller; the logits bar is 1e-3
            assert int(np.argmax(out[name])) == int(np.argmax(alone))
    finally:
        other.close()
