"""The CPU oracle (oracle/ft8_oracle.c) against fixtures generated from the unmodified reference
(oracle/make_golden.py -> tests/golden).  This is the parity pin of the oracle: everything
downstream of the spectrogram must match the reference bit for bit, the spectrogram to ~1e-9 dB."""
import hashlib
import json

import numpy as np
import pytest

from conftest import golden_case, load_golden
from ft8_demodulator_b200 import synth

SPEC_TOL_DB = 1e-9      # oracle FFT vs scipy/pocketfft, float64
SCORE_TOL = 1e-9        # scores/LLRs inherit the spectrogram difference


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def check_traced(orc, wave, fs, gold, **kw):
    """Full-path oracle decode vs a traced reference decode."""
    recs = orc.decode_all_candidates(wave, fs, **kw)
    assert [r["abs_time"] for r in recs] == gold["cand_time"].tolist()
    assert [r["abs_freq"] for r in recs] == gold["cand_freq"].tolist()
    np.testing.assert_allclose([r["score"] for r in recs], gold["cand_score"], rtol=0, atol=SCORE_TOL)
    assert [int(r["ok"]) for r in recs] == gold["ok"].tolist()
    assert [r["ldpc_errors"] for r in recs] == gold["ldpc_errors"].tolist()
    assert [r["crc_extracted"] for r in recs] == gold["crc_extracted"].tolist()
    assert [r["crc_calculated"] for r in recs] == gold["crc_calculated"].tolist()
    assert [list(r["payload"]) for r in recs] == gold["payload"].tolist()
    plain = np.array([r["plain"] for r in recs]).reshape(len(recs), 174)
    assert np.array_equal(plain, gold["plain"])
    n_llr = gold["llr"].shape[0]
    llr = np.array([r["llr"] for r in recs[:n_llr]]).reshape(n_llr, 174)
    finite = np.isfinite(gold["llr"])
    np.testing.assert_allclose(llr[finite], gold["llr"][finite], rtol=1e-9, atol=1e-9)
    # the reference's return list
    res = orc.decode_ft8_message(wave, fs, **kw)
    assert [list(r[0]) for r in res] == gold["res_payload"].tolist()
    assert [r[1] for r in res] == gold["res_hash"].tolist()
    np.testing.assert_array_equal([r[4] for r in res], gold["res_time_sec"])
    np.testing.assert_array_equal([r[5] for r in res], gold["res_freq_hz"])
    np.testing.assert_allclose([r[6] for r in res], gold["res_score"], rtol=0, atol=SCORE_TOL)
    return recs


def test_kat_encode_crc(orc):
    kat = load_golden("kat_encode.json")
    for hx, k in kat.items():
        payload = bytes.fromhex(hx)
        a91 = bytes(synth.add_crc(np.frombuffer(payload, dtype=np.uint8)))
        assert a91.hex() == k["a91"]
        assert bytes(synth.ldpc_encode(np.frombuffer(a91, dtype=np.uint8))).hex() == k["codeword"]
        assert "".join(str(int(v)) for v in synth.ft8_encode(np.frombuffer(payload, dtype=np.uint8))) == k["itones"]
        assert orc.extract_crc(a91) == k["demod_extract"] == k["crc"]
        buf = a91[:9] + bytes([a91[9] & 0xF8, 0, 0])
        assert orc.compute_crc(buf, 82) == k["demod_crc"] == k["crc"]
        bits = np.unpackbits(np.frombuffer(bytes.fromhex(k["codeword"]), dtype=np.uint8))[:174]
        assert orc.ldpc_check(bits) == k["ldpc_check"] == 0
        assert orc.pack_bits(bits, 91)[:12] == bytes.fromhex(k["a91"])
    # survey known answers
    assert kat["1c3f8a6ae207a1e39451"]["crc"] == 7997 and kat["aa0203040506070809f8"]["crc"] == 11187


def test_bp_vectors_bit_exact(orc):
    g = load_golden("bp.npz")
    for llr, it, plain, err in zip(g["llr"], g["max_iterations"], g["plain"], g["min_errors"]):
        with np.errstate(all="ignore"):
            p, e = orc.bp_decode(llr, int(it))
        assert e == err
        assert np.array_equal(p, plain)
    p, e = orc.bp_batch(g["llr"][:72], 20, n_threads=3)
    assert np.array_equal(p, g["plain"][:72]) and np.array_equal(e, g["min_errors"][:72])


def test_c1_spectrogram(orc, c1_slot):
    g = load_golden("c1.npz")
    assert sha(c1_slot.astype(np.float32)) == str(g["slot_sha256"]), "synthetic C1 slot drifted from the fixture recipe"
    mag = orc.spectrogram_pos(c1_slot, 12000, 2, 2)
    assert mag.shape == (1920, 186) and tuple(g["spec_shape"]) == (3840, 186)
    assert np.abs(mag[150:214] - g["spec_rows_150_214"]).max() < SPEC_TOL_DB
    assert np.abs(mag[::97] - g["spec_rows_stride"]).max() < SPEC_TOL_DB


@pytest.mark.parametrize("tag,kw", [
    ("k20_s1", dict(max_candidates=20, min_score=1)),
    ("k20_s3", dict(max_candidates=20, min_score=3)),
    ("k20_s10", dict(max_candidates=20, min_score=10)),
    ("k7_s2_it5", dict(max_candidates=7, min_score=2, max_iterations=5)),
    ("band_time", dict(max_candidates=20, min_score=1.5, freq_min=300.0, freq_max=1200.0, time_min=0.2, time_max=14.5)),
])
def test_c1_decode(orc, c1_slot, tag, kw):
    g = golden_case(load_golden("c1.npz"), tag)
    recs = check_traced(orc, c1_slot, 12000, g, bins_per_tone=2, steps_per_symbol=2, **kw)
    if tag == "k20_s1":   # survey known answer: (8,161) first, then the quirk's row abs_time=-20
        assert (recs[0]["abs_time"], recs[0]["abs_freq"]) == (8, 161)
        assert all(r["abs_time"] == -20 for r in recs[1:])
        assert recs[0]["payload"] == synth.C1_PAYLOAD and recs[0]["crc_calculated"] == 7997


@pytest.mark.parametrize("tag", ["k150_s4", "k38_s4", "k12_s5"])
def test_exact_score_ties(orc, c1_slot, tag):
    """Tied candidates come out in the reference's heap-array order (stable key sort, ft8_decode.py:140); a tie that
    heapq itself compares raises TypeError in the reference and in the oracle."""
    g = load_golden("ties.npz")
    kw = json.loads(str(g[tag + "__kw"]))
    if tag + "__raises" in g.files:
        with pytest.raises(TypeError):
            orc.decode_all_candidates(c1_slot, 12000, 2, 2, kw["max_candidates"], kw["min_score"], kw["max_iterations"],
                                      freq_min=kw["freq_min"], freq_max=kw["freq_max"], time_min=kw["time_min"], time_max=kw["time_max"])
        return
    gold = golden_case(g, tag)
    sc = gold["cand_score"]
    tied = np.flatnonzero(sc[1:] == sc[:-1])
    assert len(tied) >= 3
    # ... and the reference's order among them is NOT simply scan order (what a plain sort by (score, position) would give)
    scan = [(int(gold["cand_time"][i]), int(gold["cand_freq"][i])) for i in range(len(sc))]
    if tag == "k150_s4":
        assert any(scan[i] > scan[i + 1] for i in tied)
    check_traced(orc, c1_slot, 12000, gold, bins_per_tone=2, steps_per_symbol=2, **kw)


@pytest.mark.parametrize("tag", ["c12000", "c12000_band", "c6000"])
def test_complex_input(orc, tag):
    """Complex baseband into decode_ft8_message (the reference keeps the f >= 0 rows of its two-sided spectrogram,
    ft8_decode.py:316-325; src/tests/pluto-sdr/receive.py calls it that way): traced reference runs, tests/golden/cplx.npz."""
    import hashlib
    g = load_golden("cplx.npz")
    cfg, kw = json.loads(str(g[tag + "__cfg"])), json.loads(str(g[tag + "__kw"]))
    x = synth.c4_capture(**cfg)
    assert x.dtype == np.complex128 and hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == str(g[tag + "__input_sha256"])
    check_traced(orc, x, cfg["fs"], golden_case(g, tag), bins_per_tone=2, steps_per_symbol=2, **kw)


def test_wav_fixture(orc):
    g = load_golden("wav_fs20k.npz")
    x = (g["samples_int16"].astype(np.float32) / 32767).astype(np.float64)
    gold = golden_case(g, "f64")
    check_traced(orc, x, int(g["fs"]), gold, bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=10,
                 max_iterations=20, freq_max=1000)
    assert [bytes(p).hex() for p in gold["res_payload"]] == ["aa0203040506070809f8"] * 3
    # float32 reference run: same indices/payloads, scores to float32 tolerance
    g32 = golden_case(g, "f32")
    assert g32["cand_time"].tolist() == gold["cand_time"].tolist() and g32["cand_freq"].tolist() == gold["cand_freq"].tolist()
    np.testing.assert_allclose(g32["cand_score"], gold["cand_score"], rtol=1e-4)


@pytest.mark.parametrize("tag,kw", [
    ("k2000_s3", dict(max_candidates=2000, min_score=3, freq_max=3100)),
    ("k20_s3", dict(max_candidates=20, min_score=3, freq_max=3100)),
])
def test_c2_dense_band(orc, tag, kw):
    g = load_golden("c2.npz")
    slot = synth.c2_slot()[0]
    assert sha(slot.astype(np.float32)) == str(g["slot_sha256"])
    check_traced(orc, slot, 12000, golden_case(g, tag), bins_per_tone=2, steps_per_symbol=2, max_iterations=20, **kw)


@pytest.mark.parametrize("tag", ["fs6000", "fs10000_osr4", "fs3500", "fs12000_osr10", "fs32768"])
def test_other_rates(orc, tag):
    g = load_golden("rates.npz")
    r = json.loads(str(g[tag + "__recipe"]))
    pl = g[tag + "__tx_payload"]
    slot = synth.make_slot(bytes(pl), seed=r["seed"], fs=r["fs"], f0=r["f0"], start_s=r["start"], snr_db=r["snr"],
                           duration_s=r["dur"])
    assert sha(slot.astype(np.float32)) == str(g[tag + "__slot_sha256"])
    gold = golden_case(g, tag)
    check_traced(orc, slot, r["fs"], gold, max_iterations=20, **r["kw"])
    mag = orc.spectrogram_pos(slot, r["fs"], r["kw"]["bins_per_tone"], r["kw"]["steps_per_symbol"])
    stride = max(1, mag.shape[0] // 12)
    assert np.abs(mag[::stride] - g[tag + "__spec_rows_stride"]).max() < SPEC_TOL_DB
    assert bytes(pl) in [bytes(p) for p in gold["res_payload"]]


@pytest.mark.parametrize("F,T,s,b", [(300, 190, 3, 1), (101, 200, 2, 2), (64, 120, 2, 4), (400, 186, 1, 2)])
def test_drift_statement_is_pinned_by_the_plain_score_map(orc, F, T, s, b):
    """The NumPy statement of the drift-hypothesis search (an extension, no reference counterpart) with the single
    hypothesis 0 is the reference's ft8_sync_score map (ft8_decode.py:47-100) bit for bit - the C restatement of that
    map is what the golden fixtures pin; more hypotheses can only raise a score."""
    rng = np.random.default_rng(F + T)
    mag = rng.standard_normal((F, T)) * 6 - 60
    ref = orc.score_map(mag, s, b)
    got, hyp = orc.score_map_drift(mag, s, b, [0.0])
    assert np.array_equal(got, ref) and not hyp.any()
    g3, h3 = orc.score_map_drift(mag, s, b, [0.0, 0.21, -0.21])
    assert (g3 >= ref).all() and set(np.unique(h3)) <= {0, 1, 2}
    assert np.array_equal(g3[h3 == 0], ref[h3 == 0])


@pytest.mark.parametrize("F,T,s,b", [(1920, 186, 2, 2), (400, 186, 1, 2), (300, 190, 3, 1), (101, 200, 2, 2), (64, 120, 2, 4),
                                     (257, 347, 4, 4), (120, 105, 2, 2)])
def test_laplacian_statement_equals_running_sum_map(orc, F, T, s, b):
    """The path-sum Laplacian form of the score map (what csrc/sync_lap.cuh computes) against the reference-pinned running
    sum, including the time offsets whose Costas blocks are cut by the first / last column."""
    rng = np.random.default_rng(F * 7 + T)
    mag = rng.standard_normal((F, T)) * 6 - 60
    ref = orc.score_map(mag, s, b)
    got = orc.score_map_laplacian(mag, s, b)
    assert got.shape == ref.shape and ref.size
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.abs(got[fin] - ref[fin]).max() < 1e-13
    # digital silence: every comparison is exactly 0
    const = np.full((F, T), -120.00000000000001)
    z = orc.score_map_laplacian(const, s, b)
    assert np.abs(z[np.isfinite(z)]).max() < 1e-12
