#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference cannot travel to the GPU
box):  python oracle/make_golden.py [case ...]

Every case records (a) the recipe/seed of its input plus a sha256 of the exact input bytes and
(b) what the reference returned: candidate lists (order, indices, scores), per-candidate
normalised LLRs, hard decisions, ldpc_errors, CRC fields, and the decode_ft8_message result list.
Intermediate values are captured by wrapping the reference's own functions while its real
`decode_ft8_message` runs, so nothing about the orchestration is re-implemented here.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import reference_loader  # noqa: E402
from ft8_demodulator_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
R = reference_loader.load()
D = R.ft8_decode


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def traced_decode(wave, fs, llr_keep=None, **kw):
    """Run the reference decode_ft8_message, capturing its internals."""
    cap = {"cands": None, "recs": [], "bp": [], "wf_shape": None}
    orig_find, orig_dec, orig_bp = D.ft8_find_candidates, D.ft8_decode_candidate, D.bp_decode

    def find(wf, k, ms):
        cap["wf_shape"] = wf.mag.shape
        out = orig_find(wf, k, ms)
        cap["cands"] = [(c.abs_time, c.abs_freq, float(c.score)) for c in out]
        return out

    def bp(llr, it):
        plain, err = orig_bp(llr, it)
        cap["bp"].append((np.array(llr, dtype=np.float64).copy(), np.array(plain, dtype=np.uint8).copy(), int(err)))
        return plain, err

    def dec(wf, cand, it):
        ok, msg, st = orig_dec(wf, cand, it)
        cap["recs"].append((bool(ok), bytes(msg.payload), int(msg.hash), int(st.ldpc_errors),
                            int(st.crc_extracted), int(st.crc_calculated)))
        return ok, msg, st

    D.ft8_find_candidates, D.ft8_decode_candidate, D.bp_decode = find, dec, bp
    try:
        with np.errstate(all="ignore"):
            res = R.quiet(D.decode_ft8_message, wave, fs, **kw)
    finally:
        D.ft8_find_candidates, D.ft8_decode_candidate, D.bp_decode = orig_find, orig_dec, orig_bp
    n = len(cap["cands"])
    out = dict(
        wf_shape=np.array(cap["wf_shape"], dtype=np.int64),
        cand_time=np.array([c[0] for c in cap["cands"]], dtype=np.int32),
        cand_freq=np.array([c[1] for c in cap["cands"]], dtype=np.int32),
        cand_score=np.array([c[2] for c in cap["cands"]], dtype=np.float64),
        # llr: first `llr_keep` candidates only when set (keeps big-K fixtures small); plain is bit-packed
        llr=np.array([b[0] for b in cap["bp"]], dtype=np.float64).reshape(n, 174)[:llr_keep],
        plain_packed=np.packbits(np.array([b[1] for b in cap["bp"]], dtype=np.uint8).reshape(n, 174), axis=1),
        ok=np.array([r[0] for r in cap["recs"]], dtype=np.uint8),
        payload=np.array([list(r[1]) for r in cap["recs"]], dtype=np.uint8).reshape(n, 10),
        ldpc_errors=np.array([r[3] for r in cap["recs"]], dtype=np.int32),
        crc_extracted=np.array([r[4] for r in cap["recs"]], dtype=np.int32),
        crc_calculated=np.array([r[5] for r in cap["recs"]], dtype=np.int32),
        res_payload=np.array([list(bytes(m.payload)) for m, *_ in res], dtype=np.uint8).reshape(len(res), 10),
        res_hash=np.array([m.hash for m, *_ in res], dtype=np.int32),
        res_time_sec=np.array([r[2] for r in res], dtype=np.float64),
        res_freq_hz=np.array([r[3] for r in res], dtype=np.float64),
        res_score=np.array([float(r[4]) for r in res], dtype=np.float64),
    )
    return out


def save(name, **arrays):
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **arrays)
    print("  wrote %s (%.1f KB)" % (name, os.path.getsize(path) / 1024))


def prefixed(prefix, d):
    return {prefix + k: v for k, v in d.items()}


# ------------------------------------------------------------------------------------------
def case_kat():
    """Encoder known answers (reference generator: crc.py:25-47, ldpc.py:104-131, encoder.py:15-72)."""
    out = {}
    for hx in ("1c3f8a6ae207a1e39451", "1c3f8a6ae207a1e39450", "aa0203040506070809f8", "00000000000000000000",
               "fffffffffffffffffff8"):
        pl = np.frombuffer(bytes.fromhex(hx), dtype=np.uint8).copy()
        a91 = R.gen_crc.crc_generator(pl)
        cw = R.gen_ldpc.ldpc_generator(a91)
        it = R.gen_encoder.ft8_encode(pl)
        bits = np.unpackbits(cw)[:174]
        out[hx] = dict(a91=bytes(a91).hex(), crc=int(R.gen_crc.get_crc_from_a91(a91)), codeword=bytes(cw).hex(),
                       itones="".join(str(int(v)) for v in it), ldpc_check=int(R.ldpc_decoder.ldpc_check(bits)),
                       demod_crc=int(R.crc.compute_crc(bytearray(bytes(a91[:9]) + bytes([a91[9] & 0xF8, 0, 0])), 82)),
                       demod_extract=int(R.crc.extract_crc(bytearray(bytes(a91)))))
    with open(os.path.join(GOLD, "kat_encode.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("  wrote kat_encode.json")


def case_c1():
    """Config C1: one signal at -10 dB, 12 kHz, osr 2x2."""
    slot = synth.c1_slot()
    arrays = dict(slot_sha256=np.array(sha(slot.astype(np.float32))))
    S, f, t = R.quiet(R.spectrogram_analyse.calculate_spectrogram, slot, 12000, 2, 2)
    pos = S[f >= 0]
    arrays["spec_shape"] = np.array(S.shape)
    arrays["spec_rows_150_214"] = pos[150:214].copy()          # region around the 500 Hz signal
    arrays["spec_rows_stride"] = pos[::97].copy()              # sparse rows across the band
    arrays["spec_f"] = f
    arrays["spec_t"] = t
    for tag, kw in (("k20_s1", dict(max_candidates=20, min_score=1)),
                    ("k20_s3", dict(max_candidates=20, min_score=3)),
                    ("k20_s10", dict(max_candidates=20, min_score=10)),
                    ("k7_s2_it5", dict(max_candidates=7, min_score=2, max_iterations=5)),
                    ("band_time", dict(max_candidates=20, min_score=1.5, freq_min=300.0, freq_max=1200.0,
                                       time_min=0.2, time_max=14.5))):
        t0 = time.time()
        d = traced_decode(slot, 12000, bins_per_tone=2, steps_per_symbol=2, **kw)
        arrays.update(prefixed(tag + "__", d))
        print("   c1 %s: %d cands, %d decodes (%.0f s)" % (tag, len(d["cand_time"]), len(d["res_hash"]), time.time() - t0))
    save("c1.npz", **arrays)


def case_bp():
    """Standalone bp_decode vectors (ldpc_decoder.py:54-113): C5-style LLRs across SNR + edge cases."""
    llr, bits = synth.c5_llr(72, seed=11, snr2500_range=(-27.0, -16.0))
    rng = np.random.default_rng(12)
    extra = [np.zeros(174), np.full(174, np.nan), -np.abs(rng.standard_normal(174)) * 3,
             rng.standard_normal(174) * 0.01, rng.standard_normal(174) * 40.0,
             (2.0 * bits[0] - 1) * 1e-3, (2.0 * bits[1] - 1) * 50.0, np.where(bits[2] > 0, np.inf, -np.inf)]
    allv = np.concatenate([llr, np.array(extra)], axis=0)
    its = np.array([20] * len(allv) + [1, 2, 5, 40, 0, 3, 7, 11], dtype=np.int32)
    allv = np.concatenate([allv, llr[:8]], axis=0)
    plains, errs = [], []
    with np.errstate(all="ignore"):
        for v, it in zip(allv, its):
            p, e = R.ldpc_decoder.bp_decode(v.copy(), int(it))
            plains.append(np.array(p, dtype=np.uint8))
            errs.append(int(e))
    print("   bp: %d vectors, %d converge" % (len(allv), sum(1 for e in errs if e == 0)))
    save("bp.npz", llr=allv, max_iterations=its, plain=np.array(plains), min_errors=np.array(errs, dtype=np.int32))


def case_wav():
    """The reference's own fixture ft8_beacon_receiver/data/raw/ft8_fs20k_f0_550_id_1.wav, decoded the
    way from_wave.py:24-69 prepares it (int16 / 32767 -> float32)."""
    import scipy.io.wavfile as wavfile
    fs, x = wavfile.read(os.path.join(reference_loader.REFERENCE_ROOT,
                                      "src/ft8_tools/ft8_beacon_receiver/data/raw/ft8_fs20k_f0_550_id_1.wav"))
    xf = (x.astype(np.float32) / np.iinfo(np.int16).max).astype(np.float32)
    d32 = traced_decode(xf, int(fs), bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=10,
                        max_iterations=20, freq_max=1000)
    d64 = traced_decode(xf.astype(np.float64), int(fs), bins_per_tone=2, steps_per_symbol=2, max_candidates=20,
                        min_score=10, max_iterations=20, freq_max=1000)
    arrays = dict(samples_int16=x, fs=np.array(fs))
    arrays.update(prefixed("f32__", d32))
    arrays.update(prefixed("f64__", d64))
    print("   wav: f32 %d decodes, f64 %d decodes" % (len(d32["res_hash"]), len(d64["res_hash"])))
    save("wav_fs20k.npz", **arrays)


def case_c2():
    """Config C2: 50 overlapping signals, K >= #passers (quirk inactive) and K=20 (quirk active)."""
    slot, payloads, f0s, starts, snrs = synth.c2_slot()
    arrays = dict(slot_sha256=np.array(sha(slot.astype(np.float32))),
                  tx_payload=np.array([list(p) for p in payloads], dtype=np.uint8), tx_f0=np.array(f0s),
                  tx_start=np.array(starts), tx_snr=np.array(snrs))
    for tag, kw in (("k2000_s3", dict(max_candidates=2000, min_score=3, freq_max=3100)),
                    ("k20_s3", dict(max_candidates=20, min_score=3, freq_max=3100))):
        t0 = time.time()
        d = traced_decode(slot, 12000, llr_keep=48, bins_per_tone=2, steps_per_symbol=2, max_iterations=20, **kw)
        arrays.update(prefixed(tag + "__", d))
        print("   c2 %s: %d cands, %d decodes (%.0f s)" % (tag, len(d["cand_time"]), len(d["res_hash"]), time.time() - t0))
    save("c2.npz", **arrays)


def case_rates():
    """Other (fs, osr) combinations the reference's tests use."""
    arrays = {}
    recipes = {
        # test_spectrogram_analyse.py:95-123 : fs=6000, f0=0, 0 dB, osr 2x2, K=20, min_score=1
        "fs6000": dict(fs=6000, f0=0.0, snr=0.0, seed=21, dur=15.0, start=0.4,
                       kw=dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=1)),
        # test_ft8_without_channel.py:30-57 : fs=10e3 (float), f0=550, osr 4x4
        "fs10000_osr4": dict(fs=10e3, f0=550.0, snr=-12.0, seed=22, dur=14.0, start=0.6,
                             kw=dict(bins_per_tone=4, steps_per_symbol=4, max_candidates=10, min_score=4,
                                     freq_min=400.0, freq_max=800.0)),
        # test_ft8_standard.py:70-92 : odd radix nfft (fs=3500 -> nfft 1120 = 2^5*5*7)
        "fs3500": dict(fs=3500, f0=300.0, snr=-8.0, seed=23, dur=15.0, start=0.5,
                       kw=dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=20, min_score=2)),
        # README example / test_spectrogram_analyse.py:147-155 : osr 10x10 at 12 kHz (band-limited to keep it short)
        "fs12000_osr10": dict(fs=12000, f0=520.0, snr=-6.0, seed=24, dur=14.0, start=0.3,
                              kw=dict(bins_per_tone=10, steps_per_symbol=10, max_candidates=15, min_score=5,
                                      freq_min=450.0, freq_max=650.0, time_max=13.6)),
        # test_correction.py:122 : fs=32768 -> nperseg 5242, nfft 10485 = 3^2*5*233 (odd)
        "fs32768": dict(fs=32768, f0=1000.0, snr=-3.0, seed=25, dur=13.5, start=0.2,
                        kw=dict(bins_per_tone=2, steps_per_symbol=2, max_candidates=10, min_score=6,
                                freq_min=900.0, freq_max=1150.0)),
    }
    for tag, r in recipes.items():
        pl = np.random.default_rng(r["seed"]).integers(0, 256, 10, dtype=np.uint8)
        pl[9] &= 0xF8
        slot = synth.make_slot(bytes(pl), seed=r["seed"], fs=r["fs"], f0=r["f0"], start_s=r["start"], snr_db=r["snr"],
                               duration_s=r["dur"])
        t0 = time.time()
        d = traced_decode(slot, r["fs"], max_iterations=20, **r["kw"])
        arrays.update(prefixed(tag + "__", d))
        arrays[tag + "__slot_sha256"] = np.array(sha(slot.astype(np.float32)))
        arrays[tag + "__tx_payload"] = pl
        arrays[tag + "__recipe"] = np.array(json.dumps({k: v for k, v in r.items()}))
        S, f, t = R.quiet(R.spectrogram_analyse.calculate_spectrogram, slot, r["fs"], r["kw"]["bins_per_tone"],
                          r["kw"]["steps_per_symbol"])
        arrays[tag + "__spec_shape"] = np.array(S.shape)
        pos = S[f >= 0]
        arrays[tag + "__spec_rows_stride"] = pos[::max(1, pos.shape[0] // 12)].copy()
        print("   rates %s: wf %s, %d cands, %d decodes (%.0f s)" % (tag, tuple(d["wf_shape"]), len(d["cand_time"]),
                                                                    len(d["res_hash"]), time.time() - t0))
    save("rates.npz", **arrays)


C4_CASES = (("fs12000_k100", dict(fs=12000, drift_hz_s=100.0, esn0_db=40.0, seed=4)),
            ("fs12000_k3", dict(fs=12000, drift_hz_s=3.3, esn0_db=30.0, seed=5)),
            ("fs32768_k568", dict(fs=32768, drift_hz_s=568.0, esn0_db=35.0, seed=6)))
# round 2: falling chirps, another carrier offset and sample rate
C4B_CASES = (("fs12000_kneg60", dict(fs=12000, drift_hz_s=-60.0, esn0_db=35.0, seed=7, fc=2500.0)),
             ("fs12000_kneg3", dict(fs=12000, drift_hz_s=-3.3, esn0_db=32.0, seed=8, f0=700.0, fc=900.0)),
             ("fs20000_k250", dict(fs=20000, drift_hz_s=250.0, esn0_db=38.0, seed=9)))


def case_c4(cases=C4_CASES, name="c4.npz"):
    """Config C4 (satellite pass): reference correct_frequency_drift + decode of the corrected real part
    (test_correction.py:281-336 parameters), two geometries."""
    FC = R.frequency_correction
    arrays = {}
    for tag, cfg in cases:
        x = synth.c4_capture(**cfg)
        ridges, segs = [], []
        orig_spec, orig_det = FC.calculate_spectrogram, FC.detect_signal_continuity

        def spec(w, fs, b, s):
            S, f, t = orig_spec(w, fs, b, s)
            pos = S[f >= 0]
            ridges.append(np.argmax(pos, axis=0).astype(np.int32))
            return S, f, t

        def det(idx, window_size=8, max_variance=10.0):
            out = orig_det(idx, window_size=window_size, max_variance=max_variance)
            segs.append(np.array(out[0], dtype=np.int64).reshape(-1, 2))
            return out

        FC.calculate_spectrogram, FC.detect_signal_continuity = spec, det
        t0 = time.time()
        try:
            prm = {'nsync_sym': 7, 'ndata_sym': 58, 'zscore_threshold': 5, 'max_iteration_num': 400000, 'bins_per_tone': 2,
                   'steps_per_symbol': 8, 'debug_plots': False}
            with np.errstate(all="ignore"):
                corrected, rate = R.quiet(FC.correct_frequency_drift, x, cfg["fs"], 6.25, 0.16, prm)
        finally:
            FC.calculate_spectrogram, FC.detect_signal_continuity = orig_spec, orig_det
        t1 = time.time()
        d = traced_decode(np.real(corrected), cfg["fs"], llr_keep=4, bins_per_tone=2, steps_per_symbol=2, max_candidates=100,
                          min_score=6, max_iterations=40, time_min=10)
        arrays.update(prefixed(tag + "__", d))
        arrays[tag + "__cfg"] = np.array(json.dumps(cfg))
        arrays[tag + "__input_sha256"] = np.array(sha(x))
        arrays[tag + "__ridge1"] = ridges[0]
        arrays[tag + "__ridge2"] = ridges[1] if len(ridges) > 1 else np.zeros(0, np.int32)
        arrays[tag + "__segments"] = segs[0]
        arrays[tag + "__drift_per_sample"] = np.array(float(np.asarray(rate).ravel()[0]))
        arrays[tag + "__corrected_probe"] = corrected[::max(1, len(corrected) // 64)][:64]
        print("   c4 %s: drift %.6g Hz/s (true %.6g), %d decodes (correction %.0f s, decode %.0f s)" % (
            tag, float(np.asarray(rate).ravel()[0]) * cfg["fs"], cfg["drift_hz_s"], len(d["res_hash"]), t1 - t0, time.time() - t1))
    save(name, **arrays)


def case_ties():
    """Exact score ties inside the candidate list (C1 slot, time-cropped spectrogram: candidates 36 symbols apart whose
    only in-range Costas blocks are the same cells).  The reference returns tied candidates in heap-array order
    (stable key sort of the heap, ft8_decode.py:140); K = 30 adds the record replacement on top."""
    slot = synth.c1_slot()
    arrays = dict(slot_sha256=np.array(sha(slot.astype(np.float32))))
    base = dict(max_iterations=5, freq_min=300.0, freq_max=900.0, time_min=0.56, time_max=14.84)
    for tag, kw in (("k150_s4", dict(max_candidates=150, min_score=4, **base)),     # fewer passers than K
                    ("k38_s4", dict(max_candidates=38, min_score=4, **base)),       # full heap, later passers
                    ("k12_s5", dict(max_candidates=12, min_score=5, **base))):      # heapq compares a tie: TypeError
        arrays[tag + "__kw"] = np.array(json.dumps(kw))
        try:
            d = traced_decode(slot, 12000, llr_keep=2, bins_per_tone=2, steps_per_symbol=2, **kw)
        except TypeError as e:
            arrays[tag + "__raises"] = np.array("TypeError")
            print("   ties %s: the reference raises TypeError (%s)" % (tag, str(e)[:60]))
            continue
        sc = d["cand_score"]
        n_ties = int(np.sum(sc[1:] == sc[:-1]))
        arrays.update(prefixed(tag + "__", d))
        print("   ties %s: %d cands, %d adjacent exact ties, %d decodes" % (tag, len(sc), n_ties, len(d["res_hash"])))
    save("ties.npz", **arrays)


def case_cplx():
    """Complex baseband input to decode_ft8_message (the reference's SDR receive script does this,
    src/tests/pluto-sdr/receive.py:68-78): two-sided spectrogram, rows f >= 0 kept (ft8_decode.py:316-325)."""
    arrays = {}
    for tag, cfg, kw in (("c12000", dict(fs=12000, drift_hz_s=0.0, esn0_db=30.0, seed=12, f0=300.0, fc=500.0),
                          dict(max_candidates=20, min_score=6, max_iterations=20)),
                         ("c12000_band", dict(fs=12000, drift_hz_s=0.0, esn0_db=24.0, seed=13, f0=450.0, fc=900.0),
                          dict(max_candidates=60, min_score=4, max_iterations=20, freq_min=600.0, freq_max=2400.0, time_min=8.0)),
                         ("c6000", dict(fs=6000, drift_hz_s=0.0, esn0_db=28.0, seed=14, f0=200.0, fc=700.0),
                          dict(max_candidates=20, min_score=5, max_iterations=10))):
        x = synth.c4_capture(**cfg)
        t0 = time.time()
        d = traced_decode(x, cfg["fs"], llr_keep=2, bins_per_tone=2, steps_per_symbol=2, **kw)
        arrays.update(prefixed(tag + "__", d))
        arrays[tag + "__cfg"] = np.array(json.dumps(cfg))
        arrays[tag + "__kw"] = np.array(json.dumps(kw))
        arrays[tag + "__input_sha256"] = np.array(sha(x))
        print("   cplx %s: %d cands, %d decodes (%.0f s)" % (tag, len(d["cand_time"]), len(d["res_hash"]), time.time() - t0))
    save("cplx.npz", **arrays)


CASES = dict(cplx=case_cplx, ties=case_ties, c4b=lambda: case_c4(C4B_CASES, "c4b.npz"), c4=case_c4, kat=case_kat, c1=case_c1, bp=case_bp, wav=case_wav, c2=case_c2, rates=case_rates)

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or list(CASES)
    for name in which:
        print("case", name)
        CASES[name]()
