"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(ctypes) and is compared with the CPU oracle on the same seeded inputs and with the committed
golden fixtures generated from the unmodified reference.

Tolerances (all arithmetic is float64 on both sides):
  spectrogram  |dB difference| <= 1e-9   (different FFT factorisation; north_star allows 1e-4 relative)
  sync scores  |difference|    <= 1e-9   (same spectrogram in: <= 1e-12, summation order differs)
  LLRs, BP hard decisions, ldpc_errors, CRC fields, payloads, candidate indices: exact
"""
import json

import numpy as np
import pytest
import torch

from conftest import golden_case, load_golden
from ft8_demodulator_b200 import engine, synth
from ft8_demodulator_b200.ft8_demodulator import ft8_decode as fd
from ft8_demodulator_b200.ft8_demodulator import ldpc_decoder as ld
from ft8_demodulator_b200.ft8_demodulator import spectrogram_analyse as sa

pytestmark = pytest.mark.gpu

SPEC_TOL_DB = 1e-9
SCORE_TOL = 1e-9


def dev():
    return engine.require_cuda()


def gpu_mag(wave, fs, b, s, dtype=np.float64):
    nperseg, hop, nfft = fd._lib.plan(fs, b, s)
    T = fd._lib.num_frames(len(wave), nperseg, hop)
    w = engine._dev(np.asarray(wave, dtype=dtype).reshape(1, -1))
    return engine.stft_db(w, nperseg, hop, nfft, 0, (nfft + 1) // 2, 0, T)   # [1,T,F]


# ------------------------------------------------------------------------------------------ S1
@pytest.mark.parametrize("fs,b,s,n", [(12000, 2, 2, 180000), (6000, 2, 2, 90000), (10e3, 4, 4, 60000), (3500, 2, 2, 52500),
                                      (20000, 2, 2, 100000), (2000, 2, 2, 30000), (32768, 2, 2, 80000),
                                      (12000, 1, 1, 40000), (12000, 3, 5, 40000), (4500, 2, 3, 30000)])
def test_stft_vs_oracle(orc, fs, b, s, n):
    rng = np.random.default_rng(int(fs) + b + s)
    x = rng.standard_normal(n) * 0.3 + np.sin(2 * np.pi * 440.0 * np.arange(n) / fs)
    x = x.astype(np.float32).astype(np.float64)
    ref = orc.spectrogram_pos(x, fs, b, s)                       # [F,T]
    for dt in (np.float64, np.float32):
        got = gpu_mag(x, fs, b, s, dt)[0].cpu().numpy().T
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < SPEC_TOL_DB


def test_stft_crop_and_batch(orc):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 50000)).astype(np.float32)
    nperseg, hop, nfft = 1920, 960, 3840
    w = engine._dev(x)
    full = engine.stft_db(w, nperseg, hop, nfft, 0, 1921, 0, 51)
    crop = engine.stft_db(w, nperseg, hop, nfft, 100, 400, 7, 33)
    assert torch.equal(crop, full[:, 7:33, 100:400].contiguous())
    for i in range(3):
        ref = orc.spectrogram_pos(x[i].astype(np.float64), 12000, 2, 2)
        assert np.abs(full[i, :, :1920].cpu().numpy().T - ref).max() < SPEC_TOL_DB
    # Nyquist bin (row -fs/2 of the reference's two-sided layout) against a direct NumPy transform
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1920) / 1920)
    fr = np.stack([x[0, j * 960:j * 960 + 1920].astype(np.float64) * win for j in range(51)])
    nyq = 10 * np.log10(1e-12 + np.abs(np.fft.rfft(fr, n=3840)[:, 1920]) ** 2 / win.sum() ** 2)
    assert np.abs(full[0, :, 1920].cpu().numpy() - nyq).max() < 1e-8


def test_stft_fast_path_equals_generic_path(monkeypatch):
    """The specialised 15x16x8 kernel (nperseg 1920 / nfft 3840) against the generic mixed-radix kernel."""
    rng = np.random.default_rng(11)
    x = engine._dev((rng.standard_normal((2, 30000)) * 1e-3).astype(np.float64))
    fast = engine.stft_db(x, 1920, 960, 3840, 0, 1921, 0, 30)
    monkeypatch.setenv("FT8B_STFT_GENERIC", "1")
    generic = engine.stft_db(x, 1920, 960, 3840, 0, 1921, 0, 30)
    monkeypatch.delenv("FT8B_STFT_GENERIC")
    assert (fast - generic).abs().max().item() < 1e-9
    f32 = engine.stft_db(x.to(torch.float32), 1920, 960, 3840, 5, 900, 3, 17)
    ref = engine.stft_db(x.to(torch.float32).to(torch.float64), 1920, 960, 3840, 5, 900, 3, 17)
    assert torch.equal(f32, ref)


@pytest.mark.parametrize("fs,b,s,n", [(12000, 2, 2, 60000), (32768, 2, 2, 60000), (2000, 2, 2, 20000), (10000, 4, 4, 30000)])
def test_complex_input_spectrogram(fs, b, s, n):
    """Complex audio -> two-sided spectrogram (what correct_frequency_drift feeds calculate_spectrogram,
    frequency_correction.py:190-192), checked against a direct NumPy restatement of the scipy call."""
    rng = np.random.default_rng(fs + n)
    t = np.arange(n) / fs
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.1 + np.exp(2j * np.pi * (300.0 * t + 5.0 * t * t))
    nperseg, hop, nfft = fd._lib.plan(fs, b, s)
    S, f, tt = sa.calculate_spectrogram(x, fs, b, s)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nperseg) / nperseg)
    T = (n - nperseg) // hop + 1
    frames = np.stack([x[j * hop:j * hop + nperseg] * win for j in range(T)])
    ref = np.fft.fftshift(10 * np.log10(1e-12 + np.abs(np.fft.fft(frames, n=nfft, axis=1)) ** 2 / win.sum() ** 2).T, axes=0)
    assert S.shape == ref.shape == (nfft, T) and S.dtype == np.float64
    assert np.abs(S - ref).max() < 1e-8
    np.testing.assert_array_equal(f, np.fft.fftshift(np.fft.fftfreq(nfft, 1 / fs)))
    x64 = x.astype(np.complex64)                    # complex64 in -> float32 out (values widened exactly on the GPU)
    S64, _, _ = sa.calculate_spectrogram(x64, fs, b, s)
    fr64 = np.stack([x64[j * hop:j * hop + nperseg].astype(np.complex128) * win for j in range(T)])
    ref64 = np.fft.fftshift(10 * np.log10(1e-12 + np.abs(np.fft.fft(fr64, n=nfft, axis=1)) ** 2 / win.sum() ** 2).T, axes=0)
    assert S64.dtype == np.float32 and np.abs(S64 - ref64).max() < 1e-4


def test_stft_signal_floor():
    """All-zero audio sits on the 1e-12 floor, 10*log10(1e-12) (spectrogram_analyse.py:60), and the map is
    exactly constant, so every sync score of digital silence is exactly 0 as in the reference."""
    got = gpu_mag(np.zeros(20000), 12000, 2, 2)
    assert (got - 10.0 * np.log10(1e-12)).abs().max().item() < 1e-12
    assert torch.all(got == got.flatten()[0])
    scores, _ = engine.sync_scores(gpu_mag(np.zeros(180000), 12000, 2, 2), 2, 2)
    assert torch.all(scores == 0.0)
    # ... and what the search returns for silence is the reference's exact 0 (threshold on the score itself included)
    res = fd.decode_slots(np.zeros((1, 180000), dtype=np.float32), 12000, 2, 2, max_candidates=20, min_score=0, max_iterations=2)
    assert int(res.arrays.n_cand[0]) == 20 and torch.all(res.arrays.cand_score[0] == 0.0)
    assert res.arrays.cand_time[0].tolist() == [-20] * 20 and res.arrays.cand_freq[0].tolist() == list(range(20))


def test_stft_non_finite_audio_stays_non_finite():
    """inf / NaN samples: every frame that contains one comes out non-finite in every bin (never a finite garbage dB
    value the Costas search could score); frames that do not contain it are untouched."""
    rng = np.random.default_rng(5)
    x = rng.standard_normal(20000)
    ref = gpu_mag(x, 12000, 2, 2)
    for bad in (np.inf, -np.inf, np.nan):
        y = x.copy()
        y[5000] = bad                       # frames 4 and 5 (hop 960, window 1920) contain sample 5000
        got = gpu_mag(y.astype(np.float32), 12000, 2, 2)[0]
        hit = [j for j in range(got.shape[0]) if 960 * j <= 5000 < 960 * j + 1920]
        assert hit == [4, 5]
        for j in range(got.shape[0]):
            if j in hit:
                assert not torch.isfinite(got[j]).any(), (bad, j)
            else:
                assert torch.isfinite(got[j]).all()
    assert torch.isfinite(ref).all()


def test_calculate_spectrogram_dropin(c1_slot):
    g = load_golden("c1.npz")
    S, f, t = sa.calculate_spectrogram(c1_slot, 12000, 2, 2)
    assert S.shape == (3840, 186) and S.dtype == np.float64
    np.testing.assert_array_equal(f, g["spec_f"])
    np.testing.assert_allclose(t, g["spec_t"], rtol=0, atol=0)
    pos = S[f >= 0]
    assert np.abs(pos[150:214] - g["spec_rows_150_214"]).max() < SPEC_TOL_DB
    assert np.abs(pos[::97] - g["spec_rows_stride"]).max() < SPEC_TOL_DB
    # two-sided symmetry for real input: row of -f equals row of +f
    assert np.array_equal(S[1920 - 5], S[1920 + 5]) and np.array_equal(S[1], S[3839])
    S32, _, _ = sa.calculate_spectrogram(c1_slot.astype(np.float32), 12000, 2, 2)
    assert S32.dtype == np.float32 and np.abs(S32 - S).max() < 1e-4
    e = sa.calculate_spectrogram(np.zeros(10), 12000)
    assert e[0].shape == (1, 0) and len(e[1]) == 0 and len(e[2]) == 0
    sub, fsub = sa.select_frequency_band(S, f, 0, 3000)
    assert sub.shape[0] == 961 and fsub[0] == 0 and fsub[-1] == 3000


# ------------------------------------------------------------------------------------------ S2
@pytest.mark.parametrize("F,T,s,b", [(1920, 186, 2, 2), (400, 186, 1, 2), (300, 190, 3, 1), (101, 200, 2, 2),
                                     (500, 401, 4, 3), (64, 120, 2, 4), (257, 347, 4, 4)])
def test_score_map_vs_oracle(orc, F, T, s, b):
    rng = np.random.default_rng(F * 7 + T)
    mag = (rng.standard_normal((F, T)) * 6 - 60).astype(np.float64)
    ref = orc.score_map(mag, s, b)
    m = engine._dev(np.ascontiguousarray(mag.T).reshape(1, T, F))
    got, t0 = engine.sync_scores(m, s, b)
    assert got.shape[1:] == ref.shape and t0 == -10 * s
    if ref.size:
        g = got[0].cpu().numpy()
        assert np.array_equal(np.isneginf(g), np.isneginf(ref))
        fin = np.isfinite(ref)
        assert np.abs(g[fin] - ref[fin]).max() < 1e-12


@pytest.mark.parametrize("K,thr", [(20, 1.0), (20, 3.0), (20, 10.0), (1, 1.0), (7, 2.0), (300, 2.5), (4096, 1.0), (50, -100.0),
                                   (20, 100.0)])
def test_selection_vs_oracle_on_oracle_spectrogram(orc, c1_slot, K, thr):
    """Same float64 spectrogram on both sides -> identical candidate lists (order and indices)."""
    mag = orc.spectrogram_pos(c1_slot, 12000, 2, 2)
    ct, cf, cs, npass = orc.ft8_find_candidates(mag, 2, 2, K, thr)
    m = engine._dev(np.ascontiguousarray(mag.T).reshape(1, 186, 1920))
    scores, t0 = engine.sync_scores(m, 2, 2)
    gt, gf, gs, gn, gp = engine.select_candidates(scores, t0, K, thr)
    n = int(gn.cpu()[0])
    assert n == len(ct) and int(gp.cpu()[0]) == npass
    assert gt[0, :n].cpu().numpy().tolist() == ct.tolist()
    assert gf[0, :n].cpu().numpy().tolist() == cf.tolist()
    np.testing.assert_allclose(gs[0, :n].cpu().numpy(), cs, rtol=0, atol=1e-12)


@pytest.mark.parametrize("K,thr", [(20, 1.0), (20, 3.0), (1, 1.0), (300, 2.5), (4096, 1.0), (50, -100.0)])
def test_find_candidates_scores_bit_exact(orc, c1_slot, K, thr):
    """ft8b_find_candidates (selection with the spectrogram at hand, what ft8b_decode_slots uses): on the same float64
    spectrogram the returned scores are the reference's running sums BIT FOR BIT, not just within 1e-12."""
    mag = orc.spectrogram_pos(c1_slot, 12000, 2, 2)
    ct, cf, cs, npass = orc.ft8_find_candidates(mag, 2, 2, K, thr)
    m = engine._dev(np.ascontiguousarray(mag.T).reshape(1, 186, 1920))
    scores, t0 = engine.sync_scores(m, 2, 2)
    gt, gf, gs, gn, gp = engine.select_candidates(scores, t0, K, thr, mag=m, time_osr=2, freq_osr=2)
    n = int(gn.cpu()[0])
    assert n == len(ct) and int(gp.cpu()[0]) == npass
    assert gt[0, :n].cpu().numpy().tolist() == ct.tolist() and gf[0, :n].cpu().numpy().tolist() == cf.tolist()
    assert np.array_equal(gs[0, :n].cpu().numpy(), cs)


def test_find_candidates_threshold_decided_on_reference_sum(orc, c1_slot):
    """A threshold placed exactly ON a score of the map (and one ulp above / below it): the fast kernels' score for that
    cell differs from the reference's running sum in the last bits, so only the re-evaluation in the reference's
    summation order makes the pass/skip decision - and with it the whole first-K list - come out as the reference's."""
    mag = orc.spectrogram_pos(c1_slot, 12000, 2, 2)
    ref_map = orc.score_map(mag, 2, 2)                        # reference running sums, [nt, nf]
    m = engine._dev(np.ascontiguousarray(mag.T).reshape(1, 186, 1920))
    scores, t0 = engine.sync_scores(m, 2, 2)
    fast = scores[0].cpu().numpy()
    head = ref_map[0, :200]
    differing = [i for i in range(200) if np.isfinite(head[i]) and fast[0, i] != head[i] and head[i] > 0.5]
    assert differing, "expected some head scores to differ in the last bits between the two summation orders"
    for i in differing[:6]:
        for thr in (head[i], np.nextafter(head[i], np.inf), np.nextafter(head[i], -np.inf)):
            ct, cf, cs, npass = orc.ft8_find_candidates(mag, 2, 2, 20, float(thr))
            gt, gf, gs, gn, _ = engine.select_candidates(scores, t0, 20, float(thr), mag=m, time_osr=2, freq_osr=2)
            n = int(gn.cpu()[0])
            assert n == len(ct)
            assert gt[0, :n].cpu().numpy().tolist() == ct.tolist() and gf[0, :n].cpu().numpy().tolist() == cf.tolist()
            assert np.array_equal(gs[0, :n].cpu().numpy(), cs)


def test_selection_synthetic_maps():
    """Selection semantics on hand-made score maps (first-K + record replacement, K > passers, none)."""
    nt, nf = 6, 50
    rng = np.random.default_rng(0)
    base = rng.permutation(nt * nf).astype(np.float64).reshape(1, nt, nf) / 10.0   # distinct scores 0..29.9
    sc = engine._dev(base)
    flat = base.reshape(-1)
    for K, thr in [(5, 3.0), (5, 29.95), (400, 10.0), (3, 0.0), (300, -1.0)]:
        passers = [i for i in range(flat.size) if flat[i] >= thr]
        H = passers[:K]
        if len(passers) > K and H:
            best = max(H, key=lambda i: flat[i])
            for i in passers[K:]:
                if flat[i] > flat[best]:
                    H[H.index(best)] = i
                    best = i
        H.sort(key=lambda i: -flat[i])
        gt, gf, gs, gn, gp = engine.select_candidates(sc, -20, K, thr)
        n = int(gn.cpu()[0])
        assert n == len(H) and int(gp.cpu()[0]) == len(passers)
        assert (gt[0, :n].cpu().numpy() + 20).tolist() == [i // nf for i in H]
        assert gf[0, :n].cpu().numpy().tolist() == [i % nf for i in H]


# ------------------------------------------------------------------------------------------ S3
def test_bp_golden_bit_exact():
    g = load_golden("bp.npz")
    for it in np.unique(g["max_iterations"]):
        sel = g["max_iterations"] == it
        plain, err = ld.bp_decode_batch(g["llr"][sel], int(it))
        assert np.array_equal(err, g["min_errors"][sel])
        assert np.array_equal(plain, g["plain"][sel])
    p, e = ld.bp_decode(g["llr"][0], 20)
    assert np.array_equal(p, g["plain"][0]) and e == g["min_errors"][0] and p.dtype == np.uint8


def test_bp_vs_oracle_c5(orc):
    """Config C5 sample: 20k low-SNR words, hard bits / min_errors / CRC flags bit-exact vs the oracle."""
    llr, bits = synth.c5_llr(20000, seed=5)
    out = engine.ldpc_crc(engine._dev(llr), 20)
    plain, err = orc.bp_batch(llr, 20, n_threads=8)
    assert np.array_equal(out.ldpc_errors.cpu().numpy(), err)
    assert np.array_equal(out.plain.cpu().numpy(), plain)
    ok = out.ok.cpu().numpy().astype(bool)
    conv = err == 0
    assert ok.sum() > 100 and np.array_equal(ok, conv & ok)
    # converged + CRC ok words decode to the transmitted codeword's first 77 bits
    pl = out.payload.cpu().numpy()
    tx = np.packbits(bits[:, :80], axis=1)
    tx[:, 9] &= 0xF8
    assert np.array_equal(pl[ok], tx[ok][:, :10])
    assert np.array_equal(out.crc_extracted.cpu().numpy()[ok], out.crc_calculated.cpu().numpy()[ok])


def test_ldpc_check_and_normalize(orc):
    rng = np.random.default_rng(9)
    bits = rng.integers(0, 2, (64, 174), dtype=np.uint8)
    bits[0] = 0
    bits[1] = synth.codeword_bits(synth.C1_PAYLOAD)
    err = engine.ldpc_check(engine._dev(bits)).cpu().numpy()
    assert err.tolist() == [orc.ldpc_check(b) for b in bits] and err[0] == 0 and err[1] == 0
    assert ld.ldpc_check(bits[5]) == orc.ldpc_check(bits[5])
    v = rng.standard_normal((50, 174)) * 10 ** rng.uniform(-3, 3, (50, 1))
    got = engine.normalize_llr_(engine._dev(v.copy())).cpu().numpy()
    for i in range(50):
        assert np.array_equal(got[i], orc.ftx_normalize_logl(v[i]))          # bit-exact incl. numpy's pairwise mean
    a = v[0].copy()
    fd.ftx_normalize_logl(a)
    assert np.array_equal(a, got[0])


def test_llr_and_candidate_decode_on_oracle_spectrogram(orc, c1_slot):
    """Identical spectrogram in -> LLRs and every decode field bit-identical to the oracle."""
    mag = orc.spectrogram_pos(c1_slot, 12000, 2, 2)
    ct, cf, cs, _ = orc.ft8_find_candidates(mag, 2, 2, 20, 1)
    ct = np.concatenate([ct, [-20, 67, 67, -1]]).astype(np.int32)      # edges of the time range too
    cf = np.concatenate([cf, [0, 1905, 0, 700]]).astype(np.int32)
    m = engine._dev(np.ascontiguousarray(mag.T).reshape(1, 186, 1920))
    tct, tcf = engine._dev(ct.reshape(1, -1)), engine._dev(cf.reshape(1, -1))
    nc = engine._dev(np.array([len(ct)], dtype=np.int32))
    d = engine.decode_candidates(m, 2, 2, tct, tcf, nc, 20, want_llr=True, want_plain=True)
    raw = engine.extract_llr(m, 2, 2, tct, tcf, nc, normalize=False)[0].cpu().numpy()
    for i, (t, f) in enumerate(zip(ct.tolist(), cf.tolist())):
        r = orc.ft8_decode_candidate(mag, 2, 2, t, f, 20)
        assert np.array_equal(raw[i], orc.ft8_extract_likelihood(mag, 2, 2, t, f))
        assert np.array_equal(d.llr[0, i].cpu().numpy(), r["llr"])
        assert np.array_equal(d.plain[0, i].cpu().numpy(), r["plain"])
        assert int(d.ldpc_errors[0, i]) == r["ldpc_errors"] and bool(d.ok[0, i]) == r["ok"]
        assert int(d.crc_extracted[0, i]) == r["crc_extracted"] and int(d.crc_calculated[0, i]) == r["crc_calculated"]
        assert bytes(d.payload[0, i].cpu().numpy()) == r["payload"]


# ------------------------------------------------------------------------------------------ whole path
def check_against_gold(res_arrays, gold, slot=0):
    n = int(res_arrays.n_cand[slot])
    assert res_arrays.cand_time[slot, :n].tolist() == gold["cand_time"].tolist()
    assert res_arrays.cand_freq[slot, :n].tolist() == gold["cand_freq"].tolist()
    np.testing.assert_allclose(res_arrays.cand_score[slot, :n].numpy(), gold["cand_score"], rtol=0, atol=SCORE_TOL)
    assert res_arrays.ok[slot, :n].tolist() == gold["ok"].tolist()
    assert res_arrays.ldpc_errors[slot, :n].tolist() == gold["ldpc_errors"].tolist()
    assert res_arrays.crc_extracted[slot, :n].tolist() == gold["crc_extracted"].tolist()
    assert res_arrays.crc_calculated[slot, :n].tolist() == gold["crc_calculated"].tolist()
    assert res_arrays.payload[slot, :n].tolist() == gold["payload"].tolist()


def check_result_list(res, gold):
    assert [list(m.payload) for m, *_ in res] == gold["res_payload"].tolist()
    assert [m.hash for m, *_ in res] == gold["res_hash"].tolist()
    np.testing.assert_array_equal([r[2] for r in res], gold["res_time_sec"])
    np.testing.assert_array_equal([r[3] for r in res], gold["res_freq_hz"])
    np.testing.assert_allclose([r[4] for r in res], gold["res_score"], rtol=0, atol=SCORE_TOL)


@pytest.mark.parametrize("tag,kw", [
    ("k20_s1", dict(max_candidates=20, min_score=1)),
    ("k20_s3", dict(max_candidates=20, min_score=3)),
    ("k20_s10", dict(max_candidates=20, min_score=10)),
    ("k7_s2_it5", dict(max_candidates=7, min_score=2, max_iterations=5)),
    ("band_time", dict(max_candidates=20, min_score=1.5, freq_min=300.0, freq_max=1200.0, time_min=0.2, time_max=14.5)),
])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_c1_golden(c1_slot, tag, kw, dtype):
    gold = golden_case(load_golden("c1.npz"), tag)
    x = c1_slot.astype(dtype)                    # float32-representable: both dtypes carry identical values
    batch = fd.decode_slots(x.reshape(1, -1), 12000, 2, 2, **kw)
    check_against_gold(batch.arrays, gold)
    res = fd.decode_ft8_message(x, 12000, bins_per_tone=2, steps_per_symbol=2, **kw)
    check_result_list(res, gold)
    for m, st, ts, fh, sc in res:
        assert isinstance(m, fd.FT8Message) and isinstance(st, fd.FT8DecodeStatus) and len(m.payload) == 10


def test_wav_fixture_golden():
    g = load_golden("wav_fs20k.npz")
    x = (g["samples_int16"].astype(np.float32) / 32767)
    gold = golden_case(g, "f64")
    res = fd.decode_ft8_message(x, int(g["fs"]), 2, 2, 20, 10, 20, freq_max=1000)
    check_result_list(res, gold)
    assert [bytes(m.payload).hex() for m, *_ in res] == ["aa0203040506070809f8"] * 3
    assert [r[3] for r in res] == [550.0, 553.125, 546.875]


def test_pcm16_ingest(tmp_path):
    """int16 PCM normalised on the GPU == the reference's host-side float32(x)/32767 (from_wave.py:62-66)."""
    import wave
    from ft8_demodulator_b200 import from_wave
    g = load_golden("wav_fs20k.npz")
    pcm = g["samples_int16"]
    gold = golden_case(g, "f64")
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=10, max_iterations=20, freq_max=1000)
    a = fd.decode_slots(pcm.reshape(1, -1), int(g["fs"]), pcm16=True, **kw)
    b = fd.decode_slots((pcm.astype(np.float32) / 32767).reshape(1, -1), int(g["fs"]), **kw)
    for f in a.arrays.__dataclass_fields__:
        assert torch.equal(getattr(a.arrays, f), getattr(b.arrays, f)), f
    check_against_gold(a.arrays, gold)
    # spectrograms too: fast (nfft 3840) and generic kernels, odd length / unaligned start
    rng = np.random.default_rng(4)
    x16 = rng.integers(-32768, 32767, (2, 40001), dtype=np.int16)
    for fs in (12000, 6000):
        nperseg, hop, nfft = fd._lib.plan(fs, 2, 2)
        T = fd._lib.num_frames(40001, nperseg, hop)
        m16 = engine.stft_db(engine._dev(x16), nperseg, hop, nfft, 0, nfft // 2, 0, T)
        m32 = engine.stft_db(engine._dev(x16.astype(np.float32) / np.float32(32767)), nperseg, hop, nfft, 0, nfft // 2, 0, T)
        assert torch.equal(m16, m32)
    # file path
    path = str(tmp_path / "fixture.wav")
    with wave.open(path, "wb") as wf:
        wf.setnchannels(1); wf.setsampwidth(2); wf.setframerate(int(g["fs"])); wf.writeframes(pcm.tobytes())
    audio, rate = from_wave.read_wave_file(path)
    assert rate == int(g["fs"]) and audio.dtype == np.float32 and np.array_equal(audio, pcm.astype(np.float32) / 32767)
    check_result_list(from_wave.decode_ft8_from_wave(path, freq_max=1000), gold)
    # without pcm16 an int16 array is plain (unscaled) integers, as the reference would treat it: +90.3 dB everywhere
    c = fd.decode_slots(pcm.reshape(1, -1), int(g["fs"]), **kw)
    assert c.arrays.cand_freq[0, :3].tolist() == gold["cand_freq"][:3].tolist()


@pytest.mark.parametrize("tag,kw", [
    ("k2000_s3", dict(max_candidates=2000, min_score=3, freq_max=3100)),
    ("k20_s3", dict(max_candidates=20, min_score=3, freq_max=3100)),
])
def test_c2_golden(tag, kw):
    g = load_golden("c2.npz")
    slot, payloads, *_ = synth.c2_slot()
    gold = golden_case(g, tag)
    batch = fd.decode_slots(slot.astype(np.float32).reshape(1, -1), 12000, 2, 2, max_iterations=20, **kw)
    check_against_gold(batch.arrays, gold)
    if tag == "k2000_s3":
        decoded = batch.payload_sets()[0]
        assert decoded == set(bytes(p) for p in gold["res_payload"]) and len(decoded & set(payloads)) >= 20


@pytest.mark.parametrize("tag", ["c12000", "c12000_band", "c6000"])
def test_complex_input_golden(orc, tag):
    """Complex baseband through decode_slots / decode_ft8_message (full-length complex transform in S1, rows f >= 0 kept)
    against the traced reference runs of tests/golden/cplx.npz; complex64 input against the oracle on the widened values;
    NumPy, pinned and device-resident input give the same arrays."""
    g = load_golden("cplx.npz")
    cfg, kw = json.loads(str(g[tag + "__cfg"])), json.loads(str(g[tag + "__kw"]))
    x = synth.c4_capture(**cfg)
    gold = golden_case(g, tag)
    batch = fd.decode_slots(x.reshape(1, -1), cfg["fs"], 2, 2, **kw)
    check_against_gold(batch.arrays, gold)
    check_result_list(fd.decode_ft8_message(x, cfg["fs"], 2, 2, **kw), gold)
    dev_in = fd.decode_slots(torch.from_numpy(x.reshape(1, -1)).cuda(), cfg["fs"], 2, 2, **kw)
    for f in batch.arrays.__dataclass_fields__:
        assert torch.equal(getattr(batch.arrays, f), getattr(dev_in.arrays, f)), f
    x64 = x.astype(np.complex64)
    b64 = fd.decode_slots(np.stack([x64, x64[::-1].copy()]), cfg["fs"], 2, 2, **kw).arrays
    for i, w in enumerate((x64, x64[::-1].copy())):
        try:
            ref = orc.decode_all_candidates(w.astype(np.complex128), cfg["fs"], 2, 2, **kw)
        except TypeError:
            continue
        assert int(b64.n_cand[i]) == len(ref)
        for c, r in enumerate(ref):
            assert (int(b64.cand_time[i, c]), int(b64.cand_freq[i, c]), bool(b64.ok[i, c]), int(b64.ldpc_errors[i, c]),
                    bytes(b64.payload[i, c].numpy())) == (r["abs_time"], r["abs_freq"], r["ok"], r["ldpc_errors"], r["payload"])
            assert abs(float(b64.cand_score[i, c]) - r["score"]) < SCORE_TOL


@pytest.mark.parametrize("tag", ["k150_s4", "k38_s4"])
def test_exact_score_ties_golden(c1_slot, tag):
    """Exact score ties inside the list (time-cropped spectrogram): the reference's heap-array order, not scan order."""
    g = load_golden("ties.npz")
    kw = json.loads(str(g[tag + "__kw"]))
    gold = golden_case(g, tag)
    assert int(np.sum(gold["cand_score"][1:] == gold["cand_score"][:-1])) >= 3
    batch = fd.decode_slots(c1_slot.astype(np.float32).reshape(1, -1), 12000, 2, 2, **kw)
    check_against_gold(batch.arrays, gold)
    check_result_list(fd.decode_ft8_message(c1_slot, 12000, 2, 2, **kw), gold)


def test_select_ties_vs_heapq():
    """Score maps with many exact ties (quantised scores) through the stand-alone selection: wherever the reference's
    heapq does not compare a tie, the list equals what Python's own heapq gives (heap_reference_list) - K below, at and above the number of passers."""
    rng = np.random.default_rng(21)
    nt, nf, t0 = 12, 40, -4
    checked = raised = 0
    for trial in range(60):
        levels = int(rng.choice([3, 6, 12, 40]))
        smap = rng.integers(0, levels, (nt, nf)).astype(np.float64) + (rng.random((nt, nf)) < 0.5) * rng.random((nt, nf))
        K = int(rng.choice([1, 2, 5, 17, 64, 65, 200, 600]))
        thr = float(rng.choice([0.0, 1.0, levels / 2]))
        want = heap_reference_list(smap, t0, K, thr)
        ct, cf, cs, n, _ = engine.select_candidates(engine._dev(smap.reshape(1, nt, nf)), t0, K, thr)
        n = int(n[0])
        got = list(zip(ct[0, :n].tolist(), cf[0, :n].tolist(), cs[0, :n].tolist()))
        if want is None:
            raised += 1                      # the reference raises TypeError: only the multiset is defined
            continue
        checked += 1
        assert got == want, (trial, K, thr)
    assert checked >= 10 and raised >= 5


def heap_reference_list(smap, t0, K, thr):
    """ft8_decode.py:124-140 on a given score map with Python's own heapq: (abs_time, abs_freq, score) in the
    reference's order, or None where the reference raises (two equal scores compared inside heapq)."""
    import heapq

    class Cand:                              # like FT8Candidate: no ordering defined
        def __init__(self, t, f):
            self.t, self.f = t, f
    heap = []
    try:
        for ti in range(smap.shape[0]):
            for f in range(smap.shape[1]):
                score = float(smap[ti, f])
                if score == float("-inf") or score < thr:
                    continue
                if len(heap) < K:
                    heapq.heappush(heap, (-score, Cand(t0 + ti, f)))
                elif -score < heap[0][0]:
                    heapq.heapreplace(heap, (-score, Cand(t0 + ti, f)))
    except TypeError:
        return None
    return [(c.t, c.f, -k) for k, c in sorted(heap, key=lambda x: x[0])]


@pytest.mark.parametrize("tag", ["fs6000", "fs10000_osr4", "fs3500", "fs12000_osr10", "fs32768"])
def test_other_rates_golden(tag):
    g = load_golden("rates.npz")
    r = json.loads(str(g[tag + "__recipe"]))
    slot = synth.make_slot(bytes(g[tag + "__tx_payload"]), seed=r["seed"], fs=r["fs"], f0=r["f0"], start_s=r["start"],
                           snr_db=r["snr"], duration_s=r["dur"])
    gold = golden_case(g, tag)
    batch = fd.decode_slots(slot.reshape(1, -1), r["fs"], max_iterations=20, **r["kw"])
    check_against_gold(batch.arrays, gold)
    check_result_list(fd.decode_ft8_message(slot, r["fs"], max_iterations=20, **r["kw"]), gold)


def test_batch_vs_oracle_c3(orc):
    """Config C3 sample: 24 seeded slots decoded in one batch (several chunks) == oracle slot by slot."""
    n = 24
    slots, payloads = zip(*[synth.c3_slot(i) for i in range(n)])
    waves = np.stack(slots).astype(np.float32)
    batch = fd.decode_slots(waves, 12000, 2, 2, max_candidates=20, min_score=1, max_iterations=20, chunk_slots=7)
    ref = orc.decode_batch(waves.astype(np.float64), 12000, 2, 2, 20, 1, 20, n_threads=8)
    a = batch.arrays
    hits = 0
    for i in range(n):
        assert int(a.n_cand[i]) == len(ref[i])
        for c, r in enumerate(ref[i]):
            assert (int(a.cand_time[i, c]), int(a.cand_freq[i, c])) == (r["abs_time"], r["abs_freq"])
            assert abs(float(a.cand_score[i, c]) - r["score"]) < SCORE_TOL
            assert bool(a.ok[i, c]) == r["ok"] and int(a.ldpc_errors[i, c]) == r["ldpc_errors"]
            assert int(a.crc_extracted[i, c]) == r["crc_extracted"] and int(a.crc_calculated[i, c]) == r["crc_calculated"]
            assert bytes(a.payload[i, c].numpy()) == r["payload"]
        hits += payloads[i] in batch.payload_sets()[i]
    assert hits >= n // 2          # round trip: most transmitted payloads come back (SNR -16..-4 dB)
    # chunking must not change anything
    again = fd.decode_slots(waves, 12000, 2, 2, max_candidates=20, min_score=1, max_iterations=20, chunk_slots=24)
    for f in a.__dataclass_fields__:
        assert torch.equal(getattr(a, f), getattr(again.arrays, f))


def test_dropin_stage_functions(orc, c1_slot):
    """ft8_find_candidates / ft8_sync_score / ft8_decode_candidate / ft8_extract_likelihood on a host waterfall."""
    mag = orc.spectrogram_pos(c1_slot, 12000, 2, 2)
    wf = fd.create_waterfall_from_spectrogram(mag, 2, 2)
    assert wf.num_bins == 1920 and wf.num_blocks == 93
    cands = fd.ft8_find_candidates(wf, 20, 1)
    ct, cf, cs, _ = orc.ft8_find_candidates(mag, 2, 2, 20, 1)
    assert [(c.abs_time, c.abs_freq) for c in cands] == list(zip(ct.tolist(), cf.tolist()))
    assert [c.score for c in cands] == cs.tolist()                      # returned scores: the reference's own sums, bit for bit
    for c in cands[:4]:
        assert fd.ft8_sync_score(wf, c) == orc.ft8_sync_score(mag, 2, 2, c.abs_time, c.abs_freq)
    # ... also outside the range ft8_find_candidates scans (the reference function takes any position)
    from ft8_demodulator_b200.ft8_demodulator.ftx_types import FT8Candidate
    for at, af in ((-23, 7), (75, 1905), (-400, 3), (180, 100), (9, 1905)):
        assert fd.ft8_sync_score(wf, FT8Candidate(waterfall=wf, abs_time=at, abs_freq=af)) == orc.ft8_sync_score(mag, 2, 2, at, af)
    with pytest.raises(IndexError):
        fd.ft8_sync_score(wf, FT8Candidate(waterfall=wf, abs_time=0, abs_freq=1906))
    ok, msg, st = fd.ft8_decode_candidate(wf, cands[0], 20)
    assert ok and bytes(msg.payload) == synth.C1_PAYLOAD and msg.hash == 7997 and st.ldpc_errors == 0
    ok2, msg2, st2 = fd.ft8_decode_candidate(wf, cands[1], 20)
    r = orc.ft8_decode_candidate(mag, 2, 2, cands[1].abs_time, cands[1].abs_freq, 20)
    assert ok2 == r["ok"] and st2.ldpc_errors == r["ldpc_errors"] and bytes(msg2.payload) == r["payload"]
    l = np.zeros(174)
    fd.ft8_extract_likelihood(wf, cands[0], l)
    assert np.array_equal(l, orc.ft8_extract_likelihood(mag, 2, 2, cands[0].abs_time, cands[0].abs_freq))
    with pytest.raises(ValueError):
        fd.create_waterfall_from_spectrogram(np.zeros(5), 2, 2)


def test_edge_cases():
    assert fd.decode_ft8_message(np.zeros(180000), 12000) == []            # all scores 0 < 10
    assert fd.decode_ft8_message(np.zeros(1000), 12000) == []              # shorter than one window
    assert fd.decode_ft8_message(np.zeros(10), 12000) == []
    assert fd.decode_ft8_message(np.zeros(30000), 12000, min_score=-5) == []   # too few blocks: empty search range
    b = fd.decode_slots(np.zeros((0, 180000), dtype=np.float32), 12000)
    assert len(b) == 0
    with pytest.raises(NotImplementedError):
        fd.decode_slots(np.zeros((1, 180000), dtype=np.float32), 12000, max_candidates=5000, strict_topk=True)
    with pytest.raises(NotImplementedError):
        fd.decode_slots(np.zeros((1, 180000), dtype=np.float32), 12000, max_candidates=2_000_000)
    S, f, t = sa.calculate_spectrogram(np.zeros(400000), 1_000_000)         # nfft 320000: the global-memory FFT path
    assert S.shape == (320000, 4) and np.all(S == S[0, 0]) and abs(S[0, 0] + 120.0) < 1e-9
    # constant spectrogram with min_score <= 0: every score ties at 0.0 (the reference raises TypeError);
    # the deterministic tie-break returns the first K positions in scan order
    wf = fd.create_waterfall_from_spectrogram(np.zeros((120, 160)), 2, 2)
    c = fd.ft8_find_candidates(wf, 5, 0)
    assert [(x.abs_time, x.abs_freq) for x in c] == [(-20, i) for i in range(5)]


# ------------------------------------------------------------------------------------------ satellite pre-processor
@pytest.mark.parametrize("tag", ["fs12000_k3", "fs12000_k100", "fs32768_k568", "fs12000_kneg60", "fs12000_kneg3", "fs20000_k250"])
def test_drift_correction_golden(tag):
    """Config C4: correct_frequency_drift + decode of the corrected real part against the reference run recorded in
    tests/golden/c4.npz (payload-level parity per SURVEY 8d; ridge, segments and drift estimate checked as well)."""
    import hashlib
    from ft8_demodulator_b200.ft8_beacon_receiver import frequency_correction as fc
    g = load_golden("c4.npz")
    if tag + "__cfg" not in g.files:
        g = load_golden("c4b.npz")           # round 2: falling chirps, other carrier offsets, 20 kHz
    cfg = json.loads(str(g[tag + "__cfg"]))
    x = synth.c4_capture(**cfg)
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).hexdigest() == str(g[tag + "__input_sha256"])
    prm = {'nsync_sym': 7, 'ndata_sym': 58, 'zscore_threshold': 5, 'max_iteration_num': 400000, 'bins_per_tone': 2,
           'steps_per_symbol': 8, 'debug_plots': False}
    tr = {}
    corrected, rate = fc.correct_frequency_drift(x, cfg["fs"], 6.25, 0.16, prm, trace=tr)
    assert np.array_equal(tr["ridge1"], g[tag + "__ridge1"])
    assert [list(s) for s in tr["segments"]] == g[tag + "__segments"].tolist()
    assert np.mean(tr["ridge2"] == g[tag + "__ridge2"]) > 0.995        # second ridge sees a 1e-10-different waveform
    ref_rate = float(g[tag + "__drift_per_sample"])
    assert abs(rate - ref_rate) <= 1e-6 * abs(ref_rate)
    assert abs(rate * cfg["fs"] - cfg["drift_hz_s"]) < 0.05 * abs(cfg["drift_hz_s"]) + 0.2
    probe = corrected[::max(1, len(corrected) // 64)][:64]
    assert corrected.dtype == np.complex128 and np.abs(probe - g[tag + "__corrected_probe"]).max() < 1e-6
    gold = golden_case(g, tag)
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=100, min_score=6, max_iterations=40, time_min=10)
    res = fd.decode_ft8_message(np.real(corrected), cfg["fs"], **kw)
    assert [bytes(m.payload) for m, *_ in res] == [bytes(p) for p in gold["res_payload"]]
    assert len(res) >= 1 and bytes(res[0][0].payload) == synth.C1_PAYLOAD
    batch = fd.decode_slots(np.real(corrected).reshape(1, -1), cfg["fs"], **kw)
    n = int(batch.arrays.n_cand[0])
    assert batch.arrays.cand_time[0, :n].tolist() == gold["cand_time"].tolist()
    assert batch.arrays.cand_freq[0, :n].tolist() == gold["cand_freq"].tolist()


def test_column_argmax_and_dechirp_kernels():
    rng = np.random.default_rng(8)
    mag = rng.standard_normal((3, 50, 333))
    mag[1, 7, 100] = mag[1, 7, 200] = 9.0                  # tie -> first index, like np.argmax
    m = engine._dev(mag)
    idx = torch.zeros((150,), dtype=torch.int32, device=m.device)
    fd._lib.check(fd._lib.lib().ft8b_column_argmax(m.data_ptr(), 150, 333, 10, 300, idx.data_ptr(), 0))
    torch.cuda.synchronize()
    assert idx.cpu().numpy().tolist() == np.argmax(mag.reshape(150, 333)[:, 10:300], axis=1).tolist()
    from ft8_demodulator_b200.ft8_beacon_receiver import frequency_correction as fc
    n = 300000
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    c2, c3 = 568.0 / (2 * 32768.0 ** 2), -3.0 / (3 * 32768.0 ** 3)
    k = np.arange(n, dtype=np.float64)
    ref = x * np.exp(-2j * np.pi * (c2 * k ** 2 + c3 * k ** 3))
    for dt in (np.complex128, np.complex64):
        got = fc._dechirp(engine._dev(x.astype(dt)), c2, c3).cpu().numpy()
        assert np.abs(got - ref).max() < (1e-8 if dt == np.complex128 else 1e-5)


def test_sharded_decode_equals_whole_batch():
    """slots are independent, so decoding contiguous shards (what each rank of an N-GPU job does) and
    concatenating gives exactly the single-batch result; the packed records survive pack -> gather order."""
    from ft8_demodulator_b200 import dist as fdist
    w, _ = synth.synth_batch_torch(13, dev(), seed=99)
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
    whole = fd.decode_slots(w, 12000, return_device=True, **kw)
    for world in (2, 3, 8):
        parts = [fd.decode_slots(w[lo:hi], 12000, return_device=True, **kw)
                 for lo, hi in (fdist.shard_range(13, r, world) for r in range(world))]
        for f in whole.__dataclass_fields__:
            assert torch.equal(torch.cat([getattr(p, f) for p in parts]), getattr(whole, f)), (world, f)
        recs = torch.cat([fdist.pack_records(p) for p in parts])
        assert torch.equal(recs, fdist.pack_records(whole))


def test_drift_hypothesis_search_extension():
    """EXTENSION (BASELINE north_star / config 4; no counterpart in the reference, payload-level checks only):
    (1) with the single hypothesis 0 Hz/s the drift path is exactly the plain pipeline; (2) bursts drifting by
    +-3.3 Hz/s (+-50 Hz per slot) that the plain search cannot decode are recovered, with the right drift sign."""
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
    plain_in = np.stack([synth.c3_slot(i)[0] for i in range(3)])
    a = fd.decode_slots(plain_in, 12000, **kw)
    b, dr = fd.decode_slots_drift(plain_in, 12000, [0.0], **kw)
    for f in a.arrays.__dataclass_fields__:
        if f == "cand_score":       # the drift kernel sums in the reference's running order, the main kernel per cell
            assert (a.arrays.cand_score - b.arrays.cand_score).abs().max().item() < 1e-12
        else:
            assert torch.equal(getattr(a.arrays, f), getattr(b.arrays, f)), f
    assert not dr.any()
    rates = [3.3, -3.3, 2.0]
    payloads = [bytes([17 * (i + 1) % 256] * 9 + [0x50]) for i in range(3)]
    slots = np.stack([synth.drifting_slot(payloads[i], seed=50 + i, drift_hz_s=rates[i], f0=800.0 + 400 * i) for i in range(3)])
    hyps = np.arange(-4.0, 4.01, 0.25)
    plain = fd.decode_slots(slots, 12000, **kw)
    found, drift = fd.decode_slots_drift(slots, 12000, hyps, **kw)
    for i in range(3):
        assert payloads[i] not in plain.payload_sets()[i]
        assert payloads[i] in found.payload_sets()[i]
        ok = found.arrays.ok[i].numpy().astype(bool)
        pl = found.arrays.payload[i].numpy()
        est = [drift[i, c] for c in np.nonzero(ok)[0] if bytes(pl[c]) == payloads[i]]
        assert any(abs(e - rates[i]) <= 0.75 for e in est), (rates[i], est)


@pytest.mark.parametrize("F,T,s,b,hyps", [
    (1920, 186, 2, 2, list(np.linspace(-0.17, 0.17, 29))),        # C4: +-3.33 Hz/s in bins per symbol
    (961, 186, 2, 2, [0.0, 0.8, -0.8, 0.33, -0.05]),              # band-cropped, shifts up to 62 bins (widest tile)
    (300, 120, 1, 1, [0.0, 0.4, -0.4, 0.123]),                    # osr 1x1, three strips of which two are edges
    (700, 372, 4, 4, [0.0, -0.21, 0.21]),                         # osr 4x4
    (1920, 186, 2, 2, [0.0, 1.5, -1.5]),                          # shifts beyond 64 bins: the direct kernel serves the call
    (1920, 186, 2, 2, [0.0] + list(np.linspace(-0.2, 0.2, 127))),  # the full 128-hypothesis list
    (500, 186, 2, 2, [0.0, 0.45, -0.45, 0.3, -0.3, 0.7]),         # several steps inside a block (|6d| > 1.5)
])
def test_drift_sync_tiled_equals_direct(F, T, s, b, hyps):
    """The two tiled drift-hypothesis sync kernels (drift.cu: contrasts staged once per strip and walked per hypothesis;
    block sums per within-block shift pattern) against the direct one-hypothesis-at-a-time kernel on random waterfalls:
    same best hypothesis, scores within 1e-12 (the direct kernel sums in the reference's running order,
    ft8_decode.py:47-100); walk and block sums agree bit for bit (same summation tree)."""
    from ft8_demodulator_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(F + T)
    mag = (torch.randn((3, T, F), generator=g, device="cuda", dtype=torch.float64) * 12.0 - 60.0).contiguous()
    mag[1, 5, 17] = float("inf")            # non-finite cells make the scores that touch them -inf in every kernel
    mag[2, T // 2, F // 2] = float("nan")
    sb, hb, t0b = engine.sync_scores_drift(mag, s, b, hyps, variant=_lib.DRIFT_DIRECT)
    fin = torch.isfinite(sb)
    per_h = None
    got = {}
    for variant in (_lib.DRIFT_AUTO, _lib.DRIFT_WALK, _lib.DRIFT_BLOCKSUM):
        try:
            sa, ha, t0a = engine.sync_scores_drift(mag, s, b, hyps, variant=variant)
        except ValueError:
            assert variant == _lib.DRIFT_BLOCKSUM and max(abs(h) for h in hyps) > 0.5       # list outside its cover
            continue
        got[variant] = (sa, ha)
        assert t0a == t0b and sa.shape == sb.shape and sa.numel() > 0
        assert torch.equal(fin, torch.isfinite(sa))
        assert torch.equal(sa[~fin], sb[~fin])
        assert (sa[fin] - sb[fin]).abs().max().item() < 1e-12
        same = ha == hb
        if not bool(same.all()):
            # a different winner is legitimate only between hypotheses whose scores agree to rounding
            assert float((~same).float().mean()) < 1e-4
            if per_h is None:
                per_h = torch.stack([engine.sync_scores_drift(mag, s, b, [h], variant=_lib.DRIFT_DIRECT)[0] for h in hyps])
            pick = lambda hh: per_h.gather(0, hh.long()[None])[0]
            assert (pick(ha)[~same] - pick(hb)[~same]).abs().max().item() < 1e-12
    if _lib.DRIFT_BLOCKSUM in got:
        assert torch.equal(got[_lib.DRIFT_WALK][0], got[_lib.DRIFT_BLOCKSUM][0])
        assert torch.equal(got[_lib.DRIFT_WALK][1], got[_lib.DRIFT_BLOCKSUM][1])


def test_drift_sync_single_hypothesis_is_the_plain_score_map():
    """With the one hypothesis d = 0 both drift kernels return ft8b_sync_scores' map (ft8_decode.py:47-100)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    mag = (torch.randn((2, 186, 1920), generator=g, device="cuda", dtype=torch.float64) * 12.0 - 60.0).contiguous()
    ref, t0 = engine.sync_scores(mag, 2, 2)
    for variant in (0, 1, 2, 3):
        sc, hyp, t0d = engine.sync_scores_drift(mag, 2, 2, [0.0], variant=variant)
        assert t0d == t0 and not hyp.any()
        assert (sc - ref).abs().max().item() < 1e-12


def test_bp_division_matches_ieee():
    """The BP kernel's branch-free division (decode.cu div_rn_core) against IEEE division over the operand domain of
    fast_tanh / fast_atanh (ldpc_decoder.py:11-31): denominators in [60, 2.1e4], numerators of any magnitude."""
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 1 << 24
    total = 0
    for rep in range(4):
        # numerators: sign * 2^e * m with e uniform in [-1000, 20]; zeros, NaN and infinities sprinkled in
        e = torch.randint(-1000, 21, (n,), device="cuda", generator=g, dtype=torch.int64).to(torch.float64)
        m = 1.0 + torch.rand((n,), device="cuda", generator=g, dtype=torch.float64)
        sgn = torch.randint(0, 2, (n,), device="cuda", generator=g, dtype=torch.int64).to(torch.float64) * 2 - 1
        a = sgn * m * torch.exp2(e)
        a[::1001] = 0.0
        a[5::4099] = float("nan")
        a[7::8191] = float("inf")
        if rep == 0:      # the values BP actually divides: tanh / atanh numerators of O(1) arguments
            x = (torch.rand((n,), device="cuda", generator=g, dtype=torch.float64) * 2 - 1) * 4.97
            a = x * (945.0 + x * x * (105.0 + x * x))
        b = 60.0 + torch.rand((n,), device="cuda", generator=g, dtype=torch.float64) * (2.1e4 - 60.0)
        if rep == 2:
            b = 945.0 + torch.rand((n,), device="cuda", generator=g, dtype=torch.float64) * 64.0   # near tanh(0)
        bad, seen = engine.selfcheck_divide(a.contiguous(), b.contiguous())
        assert bad == 0
        assert seen > n // 2
        total += seen
    assert total > 3 * n


@pytest.mark.parametrize("b,s,F,T", [(2, 2, 1920, 186), (1, 1, 961, 93), (4, 2, 700, 190), (2, 4, 333, 371), (1, 2, 65, 120)])
def test_sync_pair_kernel_equals_single_kernel(monkeypatch, b, s, F, T):
    """Two candidates per thread (sync_pair.cuh) against one per thread (sync.cu): the same IEEE additions in the same
    order, so the score maps must be identical bit for bit - including ragged widths and -inf entries."""
    g = torch.Generator(device="cuda").manual_seed(b * 100 + s)
    mag = torch.randn((3, T, F), device="cuda", dtype=torch.float64, generator=g) * 10 - 40
    mag[1, 5, 7] = float("inf")
    mag[2, 40:44, :] = float("nan")
    pair, _ = engine.sync_scores(mag, s, b)
    monkeypatch.setenv("FT8B_SYNC_SINGLE", "1")
    single, _ = engine.sync_scores(mag, s, b)
    monkeypatch.delenv("FT8B_SYNC_SINGLE")
    assert pair.shape == single.shape and pair.numel() > 0
    assert torch.equal(pair.view(torch.int64), single.view(torch.int64))


@pytest.mark.parametrize("s,F,T", [(2, 1920, 186), (2, 961, 186), (1, 500, 93), (3, 333, 300), (2, 75, 120), (4, 130, 411),
                                   (2, 61 + 14, 110), (2, 60 + 14, 110), (2, 121 + 14, 110)])
def test_sync_laplacian_kernel_vs_walk_and_oracle(orc, monkeypatch, s, F, T):
    """freq_osr 2: the path-sum Laplacian kernel (sync_lap.cuh, FT8B_SYNC_LAP=1 - an independent formulation of the score
    map, not the default) against the difference walk (sync_pair.cuh) and against the oracle's reference-order map.  Same
    terms in a different association: 1e-12 on dB-scale waterfalls, identical -inf pattern (non-finite cells, offsets
    without comparisons); widths around the 60-candidate strips of the kernel; 93 columns per phase (unrolled tail) and
    other lengths (rolled tail).  The whole pipeline on top of it returns the same candidates and decodes."""
    g = torch.Generator(device="cuda").manual_seed(F * 3 + T + s)
    mag = torch.randn((3, T, F), device="cuda", dtype=torch.float64, generator=g) * 6 - 60
    mag[1, 5, 7] = float("inf")
    mag[2, 40:44, :] = float("nan")
    walk, t0 = engine.sync_scores(mag, s, 2)
    monkeypatch.setenv("FT8B_SYNC_LAP", "1")
    lap, _ = engine.sync_scores(mag, s, 2)
    monkeypatch.delenv("FT8B_SYNC_LAP")
    assert lap.shape == walk.shape and lap.numel() > 0
    # finite waterfall (slot 0): identical -inf pattern.  Non-finite cells (slots 1, 2): every score the reference voids
    # is void here too, plus a few whose path sums merely pass through the bad cell and cancel it algebraically
    # (inf - inf): one more reason this kernel is a cross-check, not the default.
    assert torch.equal(torch.isneginf(lap[0]), torch.isneginf(walk[0]))
    assert torch.all(torch.isneginf(lap) | ~torch.isneginf(walk))
    fin = torch.isfinite(walk) & torch.isfinite(lap)
    assert (lap[fin] - walk[fin]).abs().max().item() < 1e-12
    ref = orc.score_map(np.ascontiguousarray(mag[0].cpu().numpy().T), s, 2)
    assert np.abs(lap[0].cpu().numpy() - ref).max() < 1e-12
    # a constant waterfall (digital silence): exactly 0 wherever the three blocks and their time neighbours are complete
    const = torch.full((1, T, F), -120.00000000000001, dtype=torch.float64, device="cuda")
    monkeypatch.setenv("FT8B_SYNC_LAP", "1")
    z, _ = engine.sync_scores(const, s, 2)
    monkeypatch.delenv("FT8B_SYNC_LAP")
    nb = T // s
    rows = [u * s + r for u in range(11, max(11, nb - 69)) for r in range(s) if u * s + r < z.shape[1]]
    assert z.abs().max().item() < 1e-12 and (not rows or torch.all(z[0, rows] == 0.0))


def test_pipeline_on_laplacian_kernel_equals_default(monkeypatch):
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
    x = np.stack([synth.c3_slot(i)[0] for i in range(6)]).astype(np.float32)
    ref = fd.decode_slots(x, 12000, **kw).arrays
    monkeypatch.setenv("FT8B_SYNC_LAP", "1")
    got = fd.decode_slots(x, 12000, **kw).arrays
    monkeypatch.delenv("FT8B_SYNC_LAP")
    for f in ref.__dataclass_fields__:
        assert torch.equal(getattr(got, f), getattr(ref, f)), f


@pytest.mark.parametrize("K,thr", [(20, 1.0), (20, 4.0), (3, -5.0), (1, 1.0), (64, 2.5), (20, 50.0)])
def test_strip_summaries_equal_full_scan(monkeypatch, K, thr):
    """ft8b_decode_slots lets the sync kernel leave one arg-max per 64-frequency strip and the selection reads those
    instead of re-scanning the score map for the later record (ft8_decode.py:134-137).  Same records as the full scan,
    on noise slots (record far behind the K-th passer), on a slot whose record sits inside the first K passers, on exact
    ties of the maximum and on a threshold nothing passes."""
    rng = np.random.default_rng(K)
    slots = [synth.c3_slot(900 + i)[0] for i in range(4)]
    slots.append(rng.standard_normal(180000) * 0.1)                       # noise only
    strong = np.zeros(180000)
    w = synth.ft8_generator(synth.C1_PAYLOAD, fs=12000, f0=6.25, fc=0)     # record in the very first scan rows
    strong[:len(w) - 6000] = w[6000:]
    slots.append(strong + 1e-3 * rng.standard_normal(180000))
    slots.append(np.zeros(180000))                                        # constant spectrogram: every score ties
    waves = np.stack(slots).astype(np.float32)
    kw = dict(max_candidates=K, min_score=thr, max_iterations=5)
    fused = fd.decode_slots(waves, 12000, 2, 2, **kw).arrays
    monkeypatch.setenv("FT8B_SELECT_FULLSCAN", "1")
    scan = fd.decode_slots(waves, 12000, 2, 2, **kw).arrays
    monkeypatch.delenv("FT8B_SELECT_FULLSCAN")
    for f in fused.__dataclass_fields__:
        assert torch.equal(getattr(fused, f), getattr(scan, f)), f
    if thr < 10:
        assert int(fused.n_cand[:4].min()) == K


@pytest.mark.parametrize("fs", list(range(2000, 10500, 500)))
def test_ft8_standard_rate_sweep(orc, fs):
    """The reference's demodulator/test_ft8_standard.py:44-92, one point per sample rate: a burst at f0 = 0 in white noise
    (-10 dB, the top of its sweep) decoded with (2, 2, K=20, min_score=1, 20 iterations).  nfft = 0.32 fs runs through the
    generic mixed-radix STFT kernel with odd factors 3, 5, 7, 9, 11, 13, 17, 19; everything the reference would return
    must equal the oracle's result for the same samples, and the transmitted payload must come back."""
    rng = np.random.default_rng(fs)
    payload = bytes(rng.integers(0, 256, 10, dtype=np.uint8) & np.array([255] * 9 + [0xF8], dtype=np.uint8))
    wave = synth.ft8_generator(payload, fs=fs, f0=0, fc=0)
    noise_power = np.mean(wave ** 2) / (10 ** (-10 / 10))
    wave = wave + np.sqrt(noise_power) * rng.standard_normal(len(wave))
    ref_mag = orc.spectrogram_pos(wave, fs, 2, 2)
    got_mag = gpu_mag(wave, fs, 2, 2)[0].cpu().numpy().T
    assert got_mag.shape == ref_mag.shape and np.abs(got_mag - ref_mag).max() < SPEC_TOL_DB
    got = fd.decode_ft8_message(wave_data=wave, sample_rate=fs, bins_per_tone=2, steps_per_symbol=2, max_candidates=20,
                                min_score=1, max_iterations=20)
    ref = orc.decode_ft8_message(wave, fs, 2, 2, 20, 1, 20)
    assert len(got) == len(ref)
    for (msg, st, t_sec, f_hz, score), r in zip(got, ref):
        assert bytes(msg.payload) == r[0] and msg.hash == r[1]
        assert (st.ldpc_errors, st.crc_extracted) == (r[2], r[3])
        assert t_sec == r[4] and f_hz == r[5] and abs(score - r[6]) < SCORE_TOL
    assert payload in [bytes(m.payload) for m, *_ in got]


def test_concurrent_streams_and_threads():
    """Two host threads, each on its own CUDA stream, decode different batches at the same time (the C ABI promises
    thread safety for distinct streams: per-device tables behind mutexes, no other shared state)."""
    import threading
    d = dev()
    batches = [np.stack([synth.c3_slot(100 * k + i)[0] for i in range(6)]).astype(np.float32) for k in range(2)]
    serial = [fd.decode_slots(b, 12000, 2, 2, max_candidates=20, min_score=1, max_iterations=20).arrays for b in batches]
    out = [None, None]
    errs = []

    def work(k):
        try:
            s = torch.cuda.Stream(device=d)
            with torch.cuda.stream(s):
                for _ in range(3):
                    w = torch.from_numpy(batches[k]).to(d, non_blocking=True)
                    out[k] = fd.decode_slots(w, 12000, 2, 2, max_candidates=20, min_score=1, max_iterations=20).arrays
            s.synchronize()
        except Exception as e:          # surfaced in the main thread below
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for k in range(2):
        for f in serial[k].__dataclass_fields__:
            assert torch.equal(getattr(serial[k], f), getattr(out[k], f)), (k, f)


def _same_decodes(a, b):
    return all(torch.equal(getattr(a, f).cpu(), getattr(b, f).cpu()) for f in a.__dataclass_fields__)


def test_host_abi_path_equals_device_path():
    """ft8b_decode_slots_host (host pointers in and out, the library's own streams, double buffering and pinned
    staging) against ft8b_decode_slots on device-resident audio: pinned and pageable input, float32 / float64 / 16-bit
    PCM, strided rows, batches that are not a multiple of the piece, a single slot, and the too-short slot."""
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
    x = np.stack([synth.c3_slot(i)[0] for i in range(11)]).astype(np.float32)
    ref = fd.decode_slots(torch.from_numpy(x).cuda(), 12000, return_device=True, **kw)
    pinned = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
    pinned.copy_(torch.from_numpy(x))
    wide = np.zeros((11, x.shape[1] + 37), dtype=np.float32)
    wide[:, :x.shape[1]] = x
    for waves, piece in [(x, 128), (x, 4), (x, 1), (pinned, 4), (pinned, 11), (x.astype(np.float64), 3),
                         (torch.from_numpy(wide)[:, :x.shape[1]], 5)]:
        got = fd.decode_slots(waves, 12000, piece_slots=piece, **kw)
        assert not got.arrays.ok.is_cuda and _same_decodes(got.arrays, ref), piece
    one = fd.decode_slots(x[3], 12000, **kw)
    assert _same_decodes(one.arrays, engine.slice_slot_decodes(ref, 3, 4))
    q = np.round(x / np.abs(x).max(axis=1, keepdims=True) * 32767.0).astype(np.int16)
    ref16 = fd.decode_slots(torch.from_numpy(q).cuda(), 12000, return_device=True, pcm16=True, **kw)
    assert _same_decodes(fd.decode_slots(q, 12000, pcm16=True, piece_slots=4, **kw).arrays, ref16)
    # K = 2000 (C2-sized candidate lists) through the host path, and a band/time mask
    kw2 = dict(kw, max_candidates=2000, min_score=3, freq_max=3100.0, time_min=0.5)
    assert _same_decodes(fd.decode_slots(x[:3], 12000, piece_slots=2, **kw2).arrays,
                         fd.decode_slots(torch.from_numpy(x[:3]).cuda(), 12000, return_device=True, **kw2))
    short = fd.decode_slots(np.zeros((2, 100), dtype=np.float32), 12000, **kw)
    assert short.to_lists() == [[], []]
    # the device-side duration of the last host call is reported
    assert engine._lib.lib().ft8b_last_host_call_ms() > 0.0


def test_host_abi_path_without_torch_tensors():
    """The host entry point called the way a NumPy-only program would: ctypes + ndarray pointers, no torch objects."""
    import ctypes as C
    from ft8_demodulator_b200 import _lib
    x = np.stack([synth.c3_slot(i)[0] for i in range(3)]).astype(np.float32)
    g = fd.decode_geometry(x.shape[1], 12000, 2, 2, None, None, None, None)
    prm = engine.make_params(12000, 2, 2, 20, 1, 20, g["bin_lo"], g["bin_hi"], g["frame_lo"], g["frame_hi"])
    B, K = 3, 20
    ct, cf, nc = np.zeros((B, K), np.int32), np.zeros((B, K), np.int32), np.zeros(B, np.int32)
    cs, ok = np.zeros((B, K), np.float64), np.zeros((B, K), np.uint8)
    le, ce, cc = np.zeros((B, K), np.int32), np.zeros((B, K), np.int32), np.zeros((B, K), np.int32)
    pl = np.zeros((B, K, 10), np.uint8)
    p = lambda a: a.ctypes.data
    rc = _lib.lib().ft8b_decode_slots_host(p(x), _lib.DTYPE_F32, B, x.shape[1], x.shape[1], C.byref(prm), -1, 0, p(ct), p(cf),
                                           p(cs), p(nc), p(ok), p(le), p(ce), p(cc), p(pl))
    assert rc == 0, _lib.lib().ft8b_last_error()
    ref = fd.decode_slots(torch.from_numpy(x).cuda(), 12000, return_device=True, bins_per_tone=2, steps_per_symbol=2,
                          max_candidates=20, min_score=1, max_iterations=20).cpu()
    assert np.array_equal(ct, ref.cand_time.numpy()) and np.array_equal(cf, ref.cand_freq.numpy())
    assert np.array_equal(cs, ref.cand_score.numpy()) and np.array_equal(nc, ref.n_cand.numpy())
    assert np.array_equal(ok, ref.ok.numpy()) and np.array_equal(le, ref.ldpc_errors.numpy())
    assert np.array_equal(ce, ref.crc_extracted.numpy()) and np.array_equal(cc, ref.crc_calculated.numpy())
    assert np.array_equal(pl, ref.payload.numpy())
    # a device pointer is refused, not dereferenced on the host
    d = torch.from_numpy(x).cuda()
    rc = _lib.lib().ft8b_decode_slots_host(d.data_ptr(), _lib.DTYPE_F32, B, x.shape[1], x.shape[1], C.byref(prm), -1, 0, p(ct),
                                           p(cf), p(cs), p(nc), p(ok), p(le), p(ce), p(cc), p(pl))
    assert rc == _lib.ERR_ARG


def test_host_abi_ragged_tail_with_small_chunks():
    """A tail piece shorter than the piece can need a LARGER chunk than the full pieces did (chunk_plan balances:
    piece 8 with chunk_slots 6 runs as 2 x 4, a 5- or 6-slot tail as one chunk of 5 / 6): the workspace sized for the
    piece must serve the tail.  Also: crc_extracted / crc_calculated may be NULL on the host entry."""
    import ctypes as C
    from ft8_demodulator_b200 import _lib
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
    x = np.stack([synth.c3_slot(i)[0] for i in range(14)]).astype(np.float32)
    ref = fd.decode_slots(torch.from_numpy(x).cuda(), 12000, return_device=True, **kw)
    for n, piece, chunk in [(14, 8, 6), (13, 8, 6), (11, 8, 5), (14, 8, 3), (7, 4, 3), (14, 12, 7)]:
        got = fd.decode_slots(x[:n], 12000, piece_slots=piece, chunk_slots=chunk, **kw)
        assert _same_decodes(got.arrays, engine.slice_slot_decodes(ref, 0, n)), (n, piece, chunk)
    # NULL crc arrays through the raw host entry
    g = fd.decode_geometry(x.shape[1], 12000, 2, 2, None, None, None, None)
    prm = engine.make_params(12000, 2, 2, 20, 1, 20, g["bin_lo"], g["bin_hi"], g["frame_lo"], g["frame_hi"])
    B, K = 3, 20
    ct, cf, nc = np.zeros((B, K), np.int32), np.zeros((B, K), np.int32), np.zeros(B, np.int32)
    cs, ok, le = np.zeros((B, K), np.float64), np.zeros((B, K), np.uint8), np.zeros((B, K), np.int32)
    pl = np.zeros((B, K, 10), np.uint8)
    p = lambda a: a.ctypes.data
    rc = _lib.lib().ft8b_decode_slots_host(p(x), _lib.DTYPE_F32, B, x.shape[1], x.shape[1], C.byref(prm), -1, 0, p(ct), p(cf),
                                           p(cs), p(nc), p(ok), p(le), None, None, p(pl))
    assert rc == 0, _lib.lib().ft8b_last_error()
    r3 = engine.slice_slot_decodes(ref, 0, 3).cpu()
    assert np.array_equal(ct, r3.cand_time.numpy()) and np.array_equal(ok, r3.ok.numpy())
    assert np.array_equal(pl, r3.payload.numpy()) and np.array_equal(le, r3.ldpc_errors.numpy())


def test_host_abi_direct_result_copies():
    """Result sets of 1 MB and more go device-to-host straight into page-locked caller arrays (no mirror, no host
    memcpy): 16 slots x K = 2000 through decode_slots (which then allocates pinned result arrays) and through the raw
    entry point with pinned and with pageable NumPy-backed arrays - all equal to the device-resident run."""
    import ctypes as C
    from ft8_demodulator_b200 import _lib
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=2000, min_score=3, max_iterations=5)
    base = np.stack([synth.c3_slot(40 + i)[0] for i in range(4)]).astype(np.float32)
    x = np.concatenate([base, base[::-1], base, base[::-1]])          # 16 slots
    ref = fd.decode_slots(torch.from_numpy(x).cuda(), 12000, return_device=True, **kw)
    got = fd.decode_slots(x, 12000, piece_slots=6, **kw)
    assert got.arrays.ok.is_pinned() and _same_decodes(got.arrays, ref)
    g = fd.decode_geometry(x.shape[1], 12000, 2, 2, None, None, None, None)
    prm = engine.make_params(12000, 2, 2, 2000, 3, 5, g["bin_lo"], g["bin_hi"], g["frame_lo"], g["frame_hi"])
    for pin in (True, False):
        out = engine.alloc_slot_decodes(16, 2000, torch.device("cpu"), zero=False, pin=pin)
        for f in out.__dataclass_fields__:
            getattr(out, f).fill_(99)
        p = lambda t: t.data_ptr()
        rc = _lib.lib().ft8b_decode_slots_host(x.ctypes.data, _lib.DTYPE_F32, 16, x.shape[1], x.shape[1], C.byref(prm), -1, 5,
                                               p(out.cand_time), p(out.cand_freq), p(out.cand_score), p(out.n_cand), p(out.ok),
                                               p(out.ldpc_errors), p(out.crc_extracted), p(out.crc_calculated), p(out.payload))
        assert rc == 0, _lib.lib().ft8b_last_error()
        assert _same_decodes(out, ref), pin


def test_empty_search_defines_every_output():
    """'Nothing to search' (K = 0 is separate; here a band of 10 bins: fewer than 7 tones wide) returns zeros in all nine
    output arrays of ft8b_decode_slots, not whatever the caller's buffers held."""
    x = torch.from_numpy(np.stack([synth.c3_slot(i)[0] for i in range(2)]).astype(np.float32)).cuda()
    prm = engine.make_params(12000, 2, 2, 20, 1, 20, 100, 110, 0, 186)
    dec = engine.SlotDecoder(2, x.shape[1], prm, device=x.device)
    for f in dec.out.__dataclass_fields__:
        getattr(dec.out, f).fill_(77)
    dec.run(x)
    torch.cuda.synchronize()
    for f in dec.out.__dataclass_fields__:
        assert int(getattr(dec.out, f).abs().sum()) == 0, f


def test_non_finite_audio_is_contained():
    """Slots are independent: NaN / inf samples in one slot of a batch (a dropped SDR buffer) do not change the results of
    its neighbours, the call returns, and the poisoned slots yield no successful decode."""
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
    good = np.stack([synth.c3_slot(i)[0] for i in range(2)]).astype(np.float32)
    bad_nan, bad_inf = good[0].copy(), good[1].copy()
    bad_nan[50000:50010] = np.nan
    bad_inf[1234] = np.inf
    batch = np.stack([good[0], bad_nan, good[1], bad_inf, np.full_like(good[0], np.nan)])
    alone = fd.decode_slots(good, 12000, **kw)
    mixed = fd.decode_slots(batch, 12000, **kw)
    for f in alone.arrays.__dataclass_fields__:
        assert torch.equal(getattr(mixed.arrays, f)[[0, 2]], getattr(alone.arrays, f)), f
    assert len(alone[0]) >= 1 and mixed[0] == alone[0] and mixed[2] == alone[1]
    assert mixed[4] == []                               # all-NaN audio: nothing passes, nothing decodes
    assert not mixed.arrays.ok[4].any()


# ------------------------------------------------------------------------------------------ opt-in strict top-K
def _numpy_topk(score_map, t0, K, thr):
    """The K best passers by (score descending, scan position ascending) - plain NumPy statement of ft8b_select_topk."""
    flat = score_map.reshape(-1)
    ok = np.isfinite(flat) | (flat == np.inf)
    ok &= ~(flat < thr) & ~np.isnan(flat) & (flat != -np.inf)
    idx = np.flatnonzero(ok)
    order = idx[np.lexsort((idx, -(flat[idx] + 0.0)))][:K]
    nf = score_map.shape[1]
    return [(t0 + int(i) // nf, int(i) % nf, float(flat[i])) for i in order]


@pytest.mark.parametrize("nt,nf,K,thr,levels", [
    (88, 1906, 20, 1.0, 0), (88, 1906, 1, 1.0, 0), (88, 1906, 300, 3.0, 0), (88, 1906, 4096, -1e9, 0),
    (88, 1906, 2000, 0.0, 7),        # seven distinct values only: the tie budget decides almost everything
    (5, 40, 64, 0.5, 3),             # fewer passers than K
    (44, 947, 257, 2.0, 1000),
])
def test_strict_topk_selection_vs_numpy(nt, nf, K, thr, levels):
    """EXTENSION (opt-in): ft8b_select_topk returns the K best passers, ties by scan position."""
    g = torch.Generator(device="cuda").manual_seed(nt * nf + K)
    sc = torch.randn((4, nt, nf), generator=g, device="cuda", dtype=torch.float64) * 4.0
    if levels:
        sc = torch.round(sc * levels / 8.0) * (8.0 / levels) + 0.0
    sc[1, 0, :7] = float("-inf")
    sc[1, nt // 2, 3] = float("nan")
    sc[2, nt - 1, nf - 1] = float("inf")
    sc[3] = -0.0                                             # one value everywhere (and the sign of zero must not matter)
    sc[3, 1, 5] = 0.0
    ct, cf, cs, nc, _ = engine.select_candidates(sc.contiguous(), -20, K, thr, strict_topk=True)
    host = sc.cpu().numpy()
    for b in range(4):
        want = _numpy_topk(host[b], -20, K, thr)
        n = int(nc[b])
        assert n == len(want), (b, n, len(want))
        got = list(zip(ct[b, :n].tolist(), cf[b, :n].tolist(), cs[b, :n].tolist()))
        assert got == want, (b, [(i, x, y) for i, (x, y) in enumerate(zip(got, want)) if x != y][:3])


def test_strict_topk_pipeline(c1_slot):
    """decode_slots(strict_topk=True): the candidate list is the K best sync scores of the slot's score map,
    the default list is untouched, and a dense band decodes with K = 20."""
    kw = dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1, max_iterations=20)
    x = c1_slot.astype(np.float32).reshape(1, -1)
    ref = fd.decode_slots(x, 12000, **kw)
    strict = fd.decode_slots(x, 12000, strict_topk=True, **kw)
    a, b = ref.arrays, strict.arrays
    assert int(b.n_cand[0]) == 20
    assert (int(a.cand_time[0, 0]), int(a.cand_freq[0, 0])) == (int(b.cand_time[0, 0]), int(b.cand_freq[0, 0])) == (8, 161)
    assert (a.cand_time[0, 1:] == -20).all()                                   # the reference's list: untouched
    assert bytes(strict[0][0][0].payload) == synth.C1_PAYLOAD
    s = b.cand_score[0].numpy()
    assert (np.diff(s) <= 0).all()
    # stage by stage: same list from the stand-alone kernel on the same score map
    g = fd.decode_geometry(x.shape[1], 12000, 2, 2, None, None, None, None)
    mag = engine.stft_db(torch.from_numpy(x).cuda(), g["nperseg"], g["hop"], g["nfft"], g["bin_lo"], g["bin_hi"], g["frame_lo"], g["frame_hi"])
    sc, t0 = engine.sync_scores(mag, 2, 2)
    ct, cf, cs, nc, _ = engine.select_candidates(sc, t0, 20, 1.0, strict_topk=True)
    assert torch.equal(ct.cpu(), b.cand_time) and torch.equal(cf.cpu(), b.cand_freq) and torch.equal(cs.cpu(), b.cand_score)
    want = _numpy_topk(sc[0].cpu().numpy(), t0, 20, 1.0)
    assert list(zip(b.cand_time[0].tolist(), b.cand_freq[0].tolist(), b.cand_score[0].tolist())) == want
    # device-resident input and the one-slot drop-in call take the same flag
    dev_in = fd.decode_slots(torch.from_numpy(x).cuda(), 12000, strict_topk=True, return_device=True, **kw)
    assert torch.equal(dev_in.cand_freq.cpu(), b.cand_freq)
    assert [bytes(m.payload) for m, *_ in fd.decode_ft8_message(c1_slot, 12000, strict_topk=True, **kw)] == \
           [bytes(m.payload) for m, *_ in strict[0]]
    # the drift-hypothesis path takes the flag too: with the single hypothesis 0 Hz/s it is the plain strict pipeline
    dr, _ = fd.decode_slots_drift(x, 12000, [0.0], strict_topk=True, **kw)
    assert torch.equal(dr.arrays.cand_time, b.cand_time) and torch.equal(dr.arrays.cand_freq, b.cand_freq)
    assert torch.equal(dr.arrays.ok, b.ok) and torch.equal(dr.arrays.payload, b.payload)
    c2, payloads, *_ = synth.c2_slot()
    kw2 = dict(kw, min_score=3, freq_max=3100.0)
    assert fd.decode_ft8_message(c2, 12000, **kw2) == []                       # the reference's quirk: 20 noise cells
    found = {bytes(m.payload) for m, *_ in fd.decode_ft8_message(c2, 12000, strict_topk=True, **kw2)}
    assert len(found & {bytes(p) for p in payloads}) >= 1


@pytest.mark.parametrize("F,T,s,b,hyps", [
    (300, 190, 3, 1, [0.0, 0.4, -0.4, 0.123]),
    (101, 200, 2, 2, [0.0, 0.17, -0.17]),
    (400, 186, 1, 2, [0.0] + list(np.linspace(-0.2, 0.2, 9))),
])
def test_drift_sync_vs_cpu_statement(orc, F, T, s, b, hyps):
    """The drift-hypothesis sync kernels against the CPU statement oracle/ft8_oracle.py::score_map_drift (itself pinned
    by the reference's plain score map at the single hypothesis 0): the direct kernel performs the same additions in
    the same order, the tiled kernels the same contrasts in another association."""
    from ft8_demodulator_b200 import _lib
    rng = np.random.default_rng(F * 3 + T)
    mag = rng.standard_normal((F, T)) * 6 - 60
    ref, ref_h = orc.score_map_drift(mag, s, b, hyps)
    m = engine._dev(np.ascontiguousarray(mag.T).reshape(1, T, F))
    for variant in (_lib.DRIFT_DIRECT, _lib.DRIFT_AUTO):
        got, hyp, t0 = engine.sync_scores_drift(m, s, b, hyps, variant=variant)
        g, h = got[0].cpu().numpy(), hyp[0].cpu().numpy()
        assert g.shape == ref.shape and t0 == -10 * s
        assert np.array_equal(np.isneginf(g), np.isneginf(ref))
        fin = np.isfinite(ref)
        assert np.abs(g[fin] - ref[fin]).max() < 1e-12
        assert (h != ref_h).mean() < 1e-3          # another winner only between hypotheses that tie to rounding


@pytest.mark.parametrize("fs,b,s,kind", [(96000, 2, 2, "f32"), (192000, 2, 2, "f64"), (250000, 1, 2, "c128"), (44100, 5, 3, "f32")])
def test_high_sample_rates_beyond_the_smem_fft(orc, fs, b, s, kind):
    """Frames of more than 14 336 transform points (96 kHz audio upwards, SDR baseband) run on stft_db_big_kernel (global-
    memory Stockham): spectrogram against NumPy, the whole decode against the oracle record by record."""
    rng = np.random.default_rng(fs + b)
    dur = 14.0
    n = int(fs * dur)
    pl = bytes(rng.integers(0, 256, 9, dtype=np.uint8)) + bytes([int(rng.integers(0, 32)) << 3])
    burst = synth.ft8_generator(pl, fs, 1234.5)
    x = rng.standard_normal(n) * 0.7
    x[int(0.4 * fs):int(0.4 * fs) + len(burst)] += burst
    if kind == "c128":
        bb = synth.ft8_baseband(np.frombuffer(pl, dtype=np.uint8), fs, 1234.5)
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * 0.5
        x[int(0.4 * fs):int(0.4 * fs) + len(bb)] += bb
    elif kind == "f32":
        x = x.astype(np.float32)
    nperseg, hop, nfft = fd._lib.plan(fs, b, s)
    assert (nfft if kind == "c128" else nfft // 2) > 14336
    # spectrogram (first frames) against a NumPy restatement of the scipy call
    m = nperseg + 3 * hop
    S, f, t = sa.calculate_spectrogram(x[:m], fs, b, s)
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nperseg) / nperseg)
    xw = x[:m].astype(np.complex128 if kind == "c128" else np.float64)
    frames = np.stack([xw[j * hop:j * hop + nperseg] * win for j in range(4)])
    ref = np.fft.fftshift(10 * np.log10(1e-12 + np.abs(np.fft.fft(frames, n=nfft, axis=1)) ** 2 / win.sum() ** 2).T, axes=0)
    assert S.shape == ref.shape and np.abs(S - ref).max() < (1e-4 if kind == "f32" else 1e-8)
    # decode in a band around the burst (keeps the oracle's score map small) against the oracle
    kw = dict(max_candidates=30, min_score=5.0, max_iterations=20, freq_min=900.0, freq_max=1700.0)
    ref = orc.decode_all_candidates(x.astype(np.complex128 if kind == "c128" else np.float64), fs, b, s, **kw)
    a = fd.decode_slots(x.reshape(1, -1), fs, b, s, **kw).arrays
    assert int(a.n_cand[0]) == len(ref) and len(ref) > 0
    for c, r in enumerate(ref):
        assert (int(a.cand_time[0, c]), int(a.cand_freq[0, c]), bool(a.ok[0, c]), int(a.ldpc_errors[0, c]),
                bytes(a.payload[0, c].numpy())) == (r["abs_time"], r["abs_freq"], r["ok"], r["ldpc_errors"], r["payload"])
        assert abs(float(a.cand_score[0, c]) - r["score"]) < SCORE_TOL
    if fs != 44100:                         # (osr 5x3 at 44.1 kHz: the oracle does not recover the burst either)
        assert any(r["ok"] and r["payload"] == pl for r in ref)


@pytest.mark.parametrize("K,thr", [(5000, 2.0), (70000, 0.5)])
def test_candidate_lists_beyond_shared_memory(orc, K, thr):
    """max_candidates above 4096 (the reference takes any value; callers pass huge ones for "no limit"): the list lives in
    a global scratch instead of shared memory - fewer passers than K, more passers than K (record replacement), and the
    whole decode of two slots against the oracle."""
    slot = synth.c2_slot()[0].astype(np.float32)
    waves = np.stack([slot, slot[::-1].copy()])
    kw = dict(max_candidates=K, min_score=thr, max_iterations=3, freq_max=3100)
    a = fd.decode_slots(waves, 12000, 2, 2, **kw).arrays
    compared = 0
    for i in range(2):
        try:
            ref = orc.decode_all_candidates(waves[i].astype(np.float64), 12000, 2, 2, **kw)
        except TypeError:
            continue                            # the reference's heapq compares an exact tie in this slot and raises
        compared += 1
        assert int(a.n_cand[i]) == len(ref) and len(ref) > 4096
        got = list(zip(a.cand_time[i, :len(ref)].tolist(), a.cand_freq[i, :len(ref)].tolist()))
        assert got == [(r["abs_time"], r["abs_freq"]) for r in ref]
        assert np.abs(a.cand_score[i, :len(ref)].numpy() - np.array([r["score"] for r in ref])).max() < SCORE_TOL
        assert a.ok[i, :len(ref)].tolist() == [int(r["ok"]) for r in ref]
        assert a.ldpc_errors[i, :len(ref)].tolist() == [r["ldpc_errors"] for r in ref]
        assert [bytes(p) for p in a.payload[i, :len(ref)].numpy()] == [r["payload"] for r in ref]
    assert compared >= 1


def test_long_capture_beyond_the_walk_kernels(orc):
    """A 60-s capture (374 symbols, 325 time offsets per phase) does not fit the score sums of the walk kernels in shared
    memory (limit ~42 s): the score map then comes from sync_scores_direct_kernel, one thread per score in the
    reference's own summation order - bit-equal to the oracle's map on the same spectrogram -, and the whole decode
    equals the oracle's record by record."""
    rng = np.random.default_rng(60)
    fs, n = 12000, 60 * 12000
    x = rng.standard_normal(n) * 0.5
    payloads = []
    for k, (f0, start) in enumerate(((600.0, 3.0), (1500.0, 20.5), (2200.0, 44.0))):
        pl = bytes(rng.integers(0, 256, 9, dtype=np.uint8)) + bytes([int(rng.integers(0, 32)) << 3])
        payloads.append(pl)
        burst = synth.ft8_generator(pl, fs, f0)
        x[int(start * fs):int(start * fs) + len(burst)] += burst * 0.5
    x = x.astype(np.float32)
    mag = orc.spectrogram_pos(x.astype(np.float64), fs, 2, 2)                       # [F, T]
    dev_mag = engine._dev(np.ascontiguousarray(mag.T).reshape(1, mag.shape[1], mag.shape[0]))
    scores, t0 = engine.sync_scores(dev_mag, 2, 2)
    ref_map = orc.score_map(mag, 2, 2)
    assert scores.shape[1] == ref_map.shape[0] == 650 and t0 == -20
    assert np.array_equal(scores[0].cpu().numpy(), ref_map, equal_nan=True)
    kw = dict(max_candidates=300, min_score=6.0, max_iterations=20)
    batch = fd.decode_slots(np.stack([x, x[::-1].copy()]), fs, 2, 2, **kw)
    for i, w in enumerate((x, x[::-1].copy())):
        ref = orc.decode_all_candidates(w.astype(np.float64), fs, 2, 2, 300, 6.0, 20)
        a = batch.arrays
        assert int(a.n_cand[i]) == len(ref)
        for c, r in enumerate(ref):
            assert (int(a.cand_time[i, c]), int(a.cand_freq[i, c]), bool(a.ok[i, c]), int(a.ldpc_errors[i, c]),
                    bytes(a.payload[i, c].numpy())) == (r["abs_time"], r["abs_freq"], r["ok"], r["ldpc_errors"], r["payload"])
            assert abs(float(a.cand_score[i, c]) - r["score"]) < SCORE_TOL
    assert set(payloads) <= batch.payload_sets()[0]


# ------------------------------------------------------------------------------------------ randomised differential runs
def _tool(name):
    import importlib
    import os
    import sys
    tools = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools")
    if tools not in sys.path:
        sys.path.insert(0, tools)
    return importlib.import_module(name)


def test_fuzz_stage_functions():
    """tools/fuzz_stages.py, short run: the stage-level mirrors on arbitrary waterfalls == the oracle, everything exact."""
    out = _tool("fuzz_stages").run(80, 7)
    assert out["n_failures"] == 0, out["failures"]
    assert out["candidate_lists_compared"] >= 40 and out["candidate_decodes_compared"] >= 100


def test_fuzz_spectrogram():
    """tools/fuzz_stages.py run_spectrogram, short run: calculate_spectrogram on random geometries / lengths / input types."""
    out = _tool("fuzz_stages").run_spectrogram(40, 5)
    assert out["n_failures"] == 0, out["failures"]


def test_fuzz_whole_path():
    """tools/fuzz_parity.py, short run: random geometries / parameters / input kinds through decode_slots == the oracle."""
    out = _tool("fuzz_parity").run(14, 11)
    assert not out["failures"], out["failures"]
    assert out["candidate_records_compared"] >= 200
