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
