"""Host-side logic of the Python mirror: mask arithmetic, types, byte utilities, the signal generator,
and the multi-rank sharding helpers (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden
from ft8_demodulator_b200 import dist as fdist, synth
from ft8_demodulator_b200.ft8_demodulator import constants, crc, ft8_decode as fd, ftx_types, spectrogram_analyse as sa


@pytest.mark.parametrize("fs,b,s,n", [(12000, 2, 2, 180000), (20000, 2, 2, 252800), (10e3, 4, 4, 140000), (32768, 2, 2, 442368),
                                      (3500, 2, 2, 52500)])
@pytest.mark.parametrize("lims", [(None, None, None, None), (300.0, 1200.0, 0.2, 14.5), (None, 1000, None, None),
                                  (0, 3000, None, None), (550.0, None, 10, None), (None, None, None, 3.0),
                                  (1e9, None, None, None), (None, None, 1e9, None), (3.125, 3.125, 0.16, 0.16)])
def test_masks_match_oracle(orc, fs, b, s, n, lims):
    g = fd.decode_geometry(n, fs, b, s, *lims)
    o = orc.masks(fs, g["nperseg"], g["hop"], g["nfft"], g["T_all"], *lims)
    got = (g["bin_lo"], g["bin_hi"], g["frame_lo"], g["frame_hi"])
    if o[1] - o[0] <= 0:
        assert got[1] - got[0] == 0
    else:
        assert got[:2] == o[:2]
    if o[3] - o[2] <= 0:
        assert got[3] - got[2] == 0
    else:
        assert got[2:] == o[2:]


def test_axes_match_golden():
    g = load_golden("c1.npz")
    np.testing.assert_array_equal(sa.frequency_axis(3840, 12000), g["spec_f"])
    np.testing.assert_array_equal(sa.time_axis(180000, 1920, 960, 12000), g["spec_t"])


def test_types_and_constants():
    wf = ftx_types.FT8Waterfall(mag=np.zeros((1920, 187)), time_osr=2, freq_osr=2)
    assert wf.num_bins == 1920 and wf.num_blocks == 93
    with pytest.raises(ValueError):
        ftx_types.FT8Waterfall(mag=np.zeros(5), time_osr=2, freq_osr=2)
    wf.mag[10 + 3 * 2, 4 + 5 * 2] = 7.0
    assert ftx_types.FT8Candidate(wf, abs_time=4, abs_freq=10).get_log_power(5, 3) == 7.0
    assert bytes(ftx_types.FT8Message().payload) == bytes(10) and ftx_types.FT8DecodeStatus().ldpc_errors == 0
    assert len(constants.kFTX_LDPC_Nm) == 83 and len(constants.kFTX_LDPC_Mn) == 174
    assert sum(constants.kFTX_LDPC_Num_rows) == 522 and sorted(set(constants.kFTX_LDPC_Num_rows)) == [6, 7]
    H, G = constants.parity_check_matrix(), constants.generator_matrix()
    cw = np.concatenate([np.eye(91, dtype=np.uint8), (G @ np.eye(91, dtype=np.int64) % 2).T.astype(np.uint8)], axis=1)
    assert not ((H.astype(np.int64) @ cw.T.astype(np.int64)) % 2).any()


def test_crc_and_pack_bits_against_reference_kat():
    kat = load_golden("kat_encode.json")
    for hx, k in kat.items():
        a91 = bytearray(bytes.fromhex(k["a91"]))
        assert crc.extract_crc(a91) == k["crc"]
        buf = bytearray(a91[:10] + b"\0\0")
        buf[9] &= 0xF8
        assert crc.compute_crc(buf, 82) == k["crc"]
        out = bytearray(12)
        crc.add_crc(bytearray.fromhex(hx), out)
        assert out == a91
        bits = np.unpackbits(np.frombuffer(bytes.fromhex(k["codeword"]), dtype=np.uint8))[:174]
        assert fd.pack_bits(bits, 91) == bytearray(bytes.fromhex(k["a91"]))
    l = np.zeros(3)
    fd.ft8_extract_symbol(np.array([1., 9., 2., 3., 4., 5., 6., 7.]), l)
    assert l.tolist() == [7 - 9, 7 - 9, 9 - 5]      # s2 = wf[gray] = [1,9,3,2,5,6,4,7]


def test_generator_known_answers():
    it = synth.ft8_encode(np.frombuffer(bytes.fromhex("1c3f8a6ae207a1e39451"), dtype=np.uint8))
    assert "".join(map(str, it)) == ("3140652" "07027413236410076024143535742" "3140652"
                                     "47714340647105233673231511607" "3140652")
    w = synth.ft8_generator(np.frombuffer(synth.C1_PAYLOAD, dtype=np.uint8), 12000, 500, 0)
    assert len(w) == 151680 and abs(np.sqrt(np.mean(w ** 2)) - 0.7064115443739227) < 1e-12
    llr, bits = synth.c5_llr(32, seed=1)
    assert llr.shape == (32, 174) and np.allclose(llr.var(axis=1), 24.0)
    H = constants.parity_check_matrix().astype(np.int64)
    assert not ((H @ bits.T.astype(np.int64)) % 2).any()


def test_torch_batch_generator_decodes_with_oracle(orc):
    w, pl = synth.synth_batch_torch(3, torch.device("cpu"), seed=1234)
    recs = orc.decode_batch(w.numpy().astype(np.float64), 12000, 2, 2, 20, 1, 20, n_threads=3)
    hits = sum(any(r["ok"] and r["payload"] == bytes(pl[i]) for r in recs[i]) for i in range(3))
    assert hits >= 2


def test_shard_range():
    for n in (0, 1, 7, 4096, 10 ** 6):
        for world in (1, 2, 3, 8):
            parts = [fdist.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        fdist.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, n_total, K):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, _, w = fdist.init_process_group("gloo")
    lo, hi = fdist.shard_range(n_total, r, w)
    rec = torch.zeros((hi - lo, K, fdist.RECORD_WIDTH), dtype=torch.int32)
    rec[:, :, 4] = torch.arange(lo, hi, dtype=torch.int32)[:, None]      # slot id in the abs_time field
    rec[:, :, 5] = torch.arange(K, dtype=torch.int32)[None, :]
    allrec = fdist.gather_records(rec, n_total, r, w)
    assert allrec.shape == (n_total, K, fdist.RECORD_WIDTH)
    assert torch.equal(allrec[:, 0, 4], torch.arange(n_total, dtype=torch.int32))
    assert torch.equal(allrec[3, :, 5], torch.arange(K, dtype=torch.int32))
    dist.destroy_process_group()


def test_gather_records_world2_gloo():
    """N > 1 path on CPU: uneven shards, all_gather of the fixed-size decode records, slot order preserved."""
    mp.spawn(_gather_worker, args=(2, _free_port(), 11, 4), nprocs=2, join=True)


def test_numa_binding_helpers(monkeypatch):
    assert fdist._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert fdist._parse_cpulist("") == set()
    monkeypatch.setenv("FT8B_NUMA_BIND", "0")
    assert "skipped" in fdist.bind_host_to_gpu_numa(0)
    monkeypatch.delenv("FT8B_NUMA_BIND")
    before = os.sched_getaffinity(0)
    assert "skipped" in fdist.bind_host_to_gpu_numa(0)        # no CUDA device here: best effort, never raises
    assert os.sched_getaffinity(0) == before


def test_pack_records_layout():
    from ft8_demodulator_b200 import engine
    o = engine.alloc_slot_decodes(2, 3, torch.device("cpu"))
    o.ok[1, 2] = 1
    o.cand_time[1, 2] = -7
    o.payload[1, 2] = torch.arange(10, dtype=torch.uint8)
    rec = fdist.pack_records(o)
    assert rec.shape == (2, 3, 16) and rec[1, 2, 0] == 1 and rec[1, 2, 4] == -7 and rec[1, 2, 6:].tolist() == list(range(10))


def test_drift_host_logic_on_golden_ridge():
    """detect_signal_continuity on the reference's own ridge (tests/golden/c4.npz) reproduces its segments."""
    from ft8_demodulator_b200.ft8_beacon_receiver import frequency_correction as fc
    for tag in ("fs12000_k100", "fs12000_k3", "fs32768_k568", "fs12000_kneg60", "fs12000_kneg3", "fs20000_k250"):
        g = load_golden("c4b.npz" if tag in ("fs12000_kneg60", "fs12000_kneg3", "fs20000_k250") else "c4.npz")
        ridge = g[tag + "__ridge1"]
        n_pos = {"fs12000": 1920, "fs32768": 5243, "fs20000": 3200}[tag.split("_")[0]]
        segs, metric = fc.detect_signal_continuity(ridge, window_size=4 * 8, max_variance=0.0001 * n_pos ** 2)
        assert [list(s) for s in segs] == g[tag + "__segments"].tolist()
        assert len(metric) == len(ridge) - 32 + 1
    segs, metric = fc.detect_signal_continuity(np.arange(5), window_size=8)
    assert segs == [] and len(metric) == 5
    t = np.linspace(-1, 1, 5)
    assert abs(fc.gfsk_pulse(2.0, t)[2] - 1.0) < 1e-3 and fc.gfsk_pulse(2.0, t)[0] < 0.05
    ic, c = fc._polyfit_coefs(np.arange(12.0), 1.5 - 0.25 * np.arange(12.0) + 0.125 * np.arange(12.0) ** 2, 2)
    assert abs(ic - 1.5) < 1e-9 and np.allclose(c, [0, -0.25, 0.125])


def test_generator_mirror_names():
    """ft8_generator-compatible names reproduce the reference KATs (tests/golden/kat_encode.json)."""
    from ft8_demodulator_b200 import ft8_generator as G
    kat = load_golden("kat_encode.json")
    for hx, k in kat.items():
        pl = np.frombuffer(bytes.fromhex(hx), dtype=np.uint8)
        a91 = G.crc_generator(pl)
        assert bytes(a91).hex() == k["a91"] and int(G.get_crc_from_a91(a91)) == k["crc"]
        cw = G.ldpc_generator(a91)
        assert bytes(cw).hex() == k["codeword"]
        it = G.itones_generator(G.symbolIdSequence_generator(cw))
        assert "".join(map(str, it)) == k["itones"] == "".join(map(str, G.ft8_encode(pl)))
    w = G.ft8_generator(np.frombuffer(synth.C1_PAYLOAD, dtype=np.uint8), 12000, 500, 0)
    bb = G.ft8_modulation_waveform_generator(G.gfsk_modulation_waveform_generator(G.ft8_encode(synth.C1_PAYLOAD), 12000), 12000, 500)
    assert np.array_equal(np.real(bb), w) and np.array_equal(G.ft8_baseband_generator(synth.C1_PAYLOAD, 12000, 500), bb)


# ------------------------------------------------------------------------------------------------
# Opt-in post-processing: parity unpinned, checked by round trips
# ------------------------------------------------------------------------------------------------
def test_unpack77_round_trips():
    from ft8_demodulator_b200 import postprocess as pp
    cases = [("CQ", "K1ABC", "FN42"), ("K1ABC", "W9XYZ", "-15"), ("W9XYZ", "K1ABC", "R+05"), ("K1ABC", "W9XYZ", "RR73"),
             ("CQ DX", "G4ABC", "IO91"), ("CQ 123", "PA9XYZ", "JO22"), ("QRZ", "VK3ABC", ""), ("DE", "JA1XYZ", "73"),
             ("W1AW", "KA1ABC", "RRR"), ("K1ABC/R", "W9XYZ/R", "R EN37")]
    for a, b, g in cases:
        p = pp.pack77_standard(a, b, g)
        assert len(p) == 10 and p[9] & 7 == 0
        assert pp.unpack77(p) == " ".join(x for x in (a, b, g) if x)
    assert pp.unpack77(pp.pack77_standard("G4ABC/P", "PA9XYZ", "JO22", i3=2)) == "G4ABC/P PA9XYZ JO22"
    for text in ["TNX BOB 73 GL", "HELLO WORLD", "A", "1/2+3-4.5?"]:
        assert pp.unpack77(pp.pack77_free_text(text)) == text
    # field-level inverses over their whole ranges (sampled)
    rng = np.random.default_rng(0)
    for n in rng.integers(pp.NTOKENS + pp.MAX22, pp.NTOKENS + pp.MAX22 + 37 * 36 * 10 * 27 ** 3, 300):
        call = pp.unpack_c28(int(n))
        if len(call.replace(" ", "")) == len(call) and not call.startswith("<"):      # no embedded blanks
            assert pp.unpack_c28(pp.pack_c28(call)) == call
    for g in range(0, pp.MAXGRID4, 977):
        assert pp.pack_g15(pp.unpack_g15(g, 0)) == (g, 0)
    for rpt in range(-30, 50, 7):
        assert pp.unpack_g15(*pp.pack_g15("%+03d" % rpt)) == "%+03d" % rpt
    # telemetry and unknown types never raise
    v = (0x123456789ABCDEF012 << 6) | (5 << 3)
    assert pp.unpack77(pp.int77_to_payload(v)) == "123456789ABCDEF012"
    assert isinstance(pp.unpack77(bytes.fromhex("1c3f8a6ae207a1e39450")), str)
    with pytest.raises(ValueError):
        pp.unpack77(b"\x00" * 9)


def test_dedupe_keeps_first_occurrence():
    from ft8_demodulator_b200 import postprocess as pp
    from ft8_demodulator_b200.ft8_demodulator.ftx_types import FT8DecodeStatus, FT8Message
    mk = lambda pl, s: (FT8Message(payload=bytearray(pl), hash=1), FT8DecodeStatus(0, 1, 1), 0.0, 100.0, s)
    a, b = pp.pack77_standard("CQ", "K1ABC", "FN42"), pp.pack77_free_text("HELLO")
    res = [mk(a, 9.0), mk(b, 8.0), mk(a, 7.5), mk(a, 7.0), mk(b, 6.0)]
    out = pp.dedupe(res)
    assert [r[4] for r in out] == [9.0, 8.0]
    assert [x[0] for x in pp.annotate(res)] == ["CQ K1ABC FN42", "HELLO"]
    assert len(pp.annotate(res, unique=False)) == 5


def test_decode_params_flags_layout():
    """ft8b_decode_params as ctypes: 64 bytes, `flags` last (0 = the reference's behaviour), FT8B_FLAG_STRICT_TOPK = 1
    as the header says."""
    import ctypes as C
    import re
    from ft8_demodulator_b200 import _lib, engine
    p = engine.make_params(12000, 2, 2, 20, 1, 20, 0, 1920, 0, 186)
    assert p.flags == 0 and C.sizeof(p) == 64
    q = engine.make_params(12000, 2, 2, 20, 1, 20, 0, 1920, 0, 186, flags=_lib.FLAG_STRICT_TOPK)
    assert q.flags == 1 and _lib.DecodeParams.flags.offset == 60
    hdr = open(os.path.join(ROOT, "include", "ft8b200.h")).read()
    assert int(re.search(r"#define FT8B_FLAG_STRICT_TOPK (\d+)", hdr).group(1)) == _lib.FLAG_STRICT_TOPK


def test_batch_decodes_materialisation():
    """BatchDecodes builds the reference's 5-tuples (ft8_decode.py:383-391) from the packed arrays: successes only, in
    candidate order, candidates beyond n_cand ignored; indexing, negative indices, iteration and to_lists agree."""
    from ft8_demodulator_b200 import engine
    out = engine.alloc_slot_decodes(3, 6, torch.device("cpu"))
    out.n_cand[:] = torch.tensor([6, 2, 0])
    out.ok[0, 1] = 1; out.ok[0, 4] = 1; out.ok[1, 0] = 1
    out.ok[1, 3] = 1                                   # beyond n_cand[1] = 2: not a candidate of this slot
    out.payload[0, 1] = torch.arange(10, dtype=torch.uint8)
    out.payload[0, 4] = torch.arange(10, 20, dtype=torch.uint8)
    out.crc_calculated[0, 1] = 7997; out.crc_extracted[0, 1] = 7997; out.ldpc_errors[0, 4] = 0
    out.cand_time[0, 1] = 8; out.cand_freq[0, 1] = 161; out.cand_score[0, 1] = 10.5
    out.cand_time[0, 4] = -20; out.cand_freq[0, 4] = 30; out.cand_score[0, 4] = 2.25
    b = fd.BatchDecodes(out, 12000, 2)
    assert len(b) == 3
    r0 = b[0]
    assert [bytes(m.payload) for m, *_ in r0] == [bytes(range(10)), bytes(range(10, 20))]
    m, st, t, f, sc = r0[0]
    assert isinstance(m, ftx_types.FT8Message) and isinstance(m.payload, bytearray) and m.hash == 7997
    assert isinstance(st, ftx_types.FT8DecodeStatus) and (st.ldpc_errors, st.crc_extracted, st.crc_calculated) == (0, 7997, 7997)
    assert (t, f, sc) == (8 / 12000, 161 / 2 * 6.25, 10.5)
    assert r0[1][2:] == (-20 / 12000, 30 / 2 * 6.25, 2.25)
    assert len(b[1]) == 1 and b[2] == [] and b[-1] == [] and b[-3] == r0
    assert [len(x) for x in b] == [2, 1, 0] and b.to_lists() == [b[0], b[1], b[2]]
    with pytest.raises(IndexError):
        b[3]
    assert b.payload_sets()[0] == {bytes(range(10)), bytes(range(10, 20))}
