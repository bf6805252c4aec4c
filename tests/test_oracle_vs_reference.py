"""Oracle and generator against the LIVE reference (only where /root/reference exists)."""
import numpy as np
import pytest

from ft8_demodulator_b200 import synth
from ft8_demodulator_b200.ft8_demodulator import constants

pytestmark = pytest.mark.needs_reference


@pytest.fixture(scope="module")
def ref():
    import reference_loader
    return reference_loader.load()


def test_tables_and_generator_matrix(ref):
    assert constants.kFTX_LDPC_Nm == [list(r) for r in ref.constants.kFTX_LDPC_Nm]
    assert constants.kFTX_LDPC_Mn == [list(r) for r in ref.constants.kFTX_LDPC_Mn]
    assert constants.kFTX_LDPC_Num_rows == list(ref.constants.kFTX_LDPC_Num_rows)
    assert np.array_equal(constants.generator_matrix(), np.unpackbits(ref.gen_ldpc.FT8_LDPC_G, axis=1)[:, :91])


def test_waveform_bit_identical(ref):
    pl = np.frombuffer(bytes.fromhex("aa0203040506070809f8"), dtype=np.uint8).copy()
    w = ref.quiet(ref.generator.ft8_generator, pl, 6000, 321.5, 0)
    assert np.array_equal(w, synth.ft8_generator(pl, 6000, 321.5, 0))


def test_stages_bit_exact_on_reference_spectrogram(ref, orc):
    rng = np.random.default_rng(77)
    pl = rng.integers(0, 256, 10, dtype=np.uint8)
    slot = synth.make_slot(bytes(pl), seed=77, fs=4000, f0=700.0, start_s=0.3, snr_db=-9.0, duration_s=15.0)
    S, f, t = ref.quiet(ref.spectrogram_analyse.calculate_spectrogram, slot, 4000, 2, 2)
    pos = np.ascontiguousarray(S[f >= 0])
    assert np.abs(orc.spectrogram_pos(slot, 4000, 2, 2) - pos).max() < 1e-9
    wf = ref.ft8_decode.create_waterfall_from_spectrogram(pos, 2, 2)
    cands = ref.quiet(ref.ft8_decode.ft8_find_candidates, wf, 12, 2)
    ct, cf, cs, _ = orc.ft8_find_candidates(pos, 2, 2, 12, 2)
    assert [(c.abs_time, c.abs_freq, c.score) for c in cands] == list(zip(ct.tolist(), cf.tolist(), cs.tolist()))
    for c in cands[:6]:
        l = np.zeros(174)
        ref.ft8_decode.ft8_extract_likelihood(wf, c, l)
        ref.ft8_decode.ftx_normalize_logl(l)
        r = orc.ft8_decode_candidate(pos, 2, 2, c.abs_time, c.abs_freq, 20)
        assert np.array_equal(l, r["llr"])
        ok, msg, st = ref.ft8_decode.ft8_decode_candidate(wf, c, 20)
        assert (ok, bytes(msg.payload), st.ldpc_errors, st.crc_extracted, st.crc_calculated) == \
               (r["ok"], r["payload"], r["ldpc_errors"], r["crc_extracted"], r["crc_calculated"])


def test_drift_host_logic_vs_reference(ref):
    """The closed-form continuity metric / fits against the reference's scikit-learn based ones."""
    from ft8_demodulator_b200.ft8_beacon_receiver import frequency_correction as fc
    FC = ref.frequency_correction
    if FC is None:
        pytest.skip("scikit-learn not available: reference frequency_correction not importable")
    rng = np.random.default_rng(3)
    ridge = np.concatenate([rng.integers(0, 1900, 150), np.round(400 + 0.8 * np.arange(300) + rng.normal(0, 0.6, 300)).astype(int),
                            rng.integers(0, 1900, 120)])
    a = fc.detect_signal_continuity(ridge, window_size=32, max_variance=0.0001 * 1920 ** 2)
    b = ref.quiet(FC.detect_signal_continuity, ridge, window_size=32, max_variance=0.0001 * 1920 ** 2)
    assert [tuple(s) for s in a[0]] == [tuple(s) for s in b[0]]
    np.testing.assert_allclose(a[1], b[1], rtol=1e-9, atol=1e-9)
    t = np.linspace(-1, 1, 9)
    np.testing.assert_array_equal(fc.gfsk_pulse(2.0, t), FC.gfsk_pulse(2.0, t))
    from sklearn.linear_model import LinearRegression
    from sklearn.preprocessing import PolynomialFeatures
    x = np.sort(rng.uniform(10, 25, 40)); y = 3.0 + 0.7 * x - 0.02 * x ** 2 + rng.normal(0, 0.1, 40)
    for deg in (1, 2):
        m = LinearRegression().fit(PolynomialFeatures(degree=deg).fit_transform(x.reshape(-1, 1)), y)
        ic, c = fc._polyfit_coefs(x, y, deg)
        assert abs(ic - m.intercept_) < 1e-8 and np.allclose(c[1:], m.coef_[1:], rtol=1e-8)
