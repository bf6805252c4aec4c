"""(174,91) LDPC Tanner-graph tables of the FT8 protocol.

Mirrors the names the reference exposes in
`ft8_tools/ft8_demodulator/constants.py:13-283` (`kFTX_LDPC_Nm`, `kFTX_LDPC_Mn`,
`kFTX_LDPC_Num_rows`, `FTX_LDPC_M`, `FTX_LDPC_N`) so callers can switch imports.

Only the 83 parity checks are stored (protocol data: for each check the 1-based
codeword bit indices that must XOR to zero, 0 = padding, three checks per text
line).  Everything else is DERIVED here: the per-bit check lists `kFTX_LDPC_Mn`
(ascending check order, which equals the reference table row for row), the row
weights, and the systematic generator used by the synthetic-signal encoder
(parity = B^-1 A msg for H = [A | B]).
"""
import numpy as np

FTX_LDPC_M = 83   # parity checks
FTX_LDPC_N = 174  # codeword bits
FTX_LDPC_K = 91   # message bits (77 payload + 14 CRC)

_NM_TEXT = """
      4  31  59  91  92  96 153   5  32  60  93 115 146   0   6  24  61  94 122 151   0
      7  33  62  95  96 143   0   8  25  63  83  93  96 148   6  32  64  97 126 138   0
      5  34  65  78  98 107 154   9  35  66  99 139 146   0  10  36  67 100 107 126   0
     11  37  67  87 101 139 158  12  38  68 102 105 155   0  13  39  69 103 149 162   0
      8  40  70  82 104 114 145  14  41  71  88 102 123 156  15  42  59 106 123 159   0
      1  33  72 106 107 157   0  16  43  73 108 141 160   0  17  37  74  81 109 131 154
     11  44  75 110 121 166   0  45  55  64 111 130 161 173   8  46  71 112 119 166   0
     18  36  76  89 113 114 143  19  38  77 104 116 163   0  20  47  70  92 138 165   0
      2  48  74 113 128 160   0  21  45  78  83 117 121 151  22  47  58 118 127 164   0
     16  39  62 112 134 158   0  23  43  79 120 131 145   0  19  35  59  73 110 125 161
     20  36  63  94 136 161   0  14  31  79  98 132 164   0   3  44  80 124 127 169   0
     19  46  81 117 135 167   0   7  49  58  90 100 105 168  12  50  61 118 119 144   0
     13  51  64 114 118 157   0  24  52  76 129 148 149   0  25  53  69  90 101 130 156
     20  46  65  80 120 140 170  21  54  77 100 140 171   0  35  82 133 142 171 174   0
     14  30  83 113 125 170   0   4  29  68 120 134 173   0   1   4  52  57  86 136 152
     26  51  56  91 122 137 168  52  84 110 115 145 168   0   7  50  81  99 132 173   0
     23  55  67  95 172 174   0  26  41  77 109 141 148   0   2  27  41  61  62 115 133
     27  40  56 124 125 126   0  18  49  55 124 141 167   0   6  33  85 108 116 156   0
     28  48  70  85 105 129 158   9  54  63 131 147 155   0  22  53  68 109 121 174   0
      3  13  48  78  95 123   0  31  69 133 150 155 169   0  12  43  66  89  97 135 159
      5  39  75 102 136 167   0   2  54  86 101 135 164   0  15  56  87 108 119 171   0
     10  44  82  91 111 144 149  23  34  71  94 127 153   0  11  49  88  92 142 157   0
     29  34  87  97 147 162   0  30  50  60  86 137 142 162  10  53  66  84 112 128 165
     22  57  85  93 140 159   0  28  32  72 103 132 166   0  28  29  84  88 117 143 150
      1  26  45  80 128 147   0  17  27  89 103 116 153   0  51  57  98 163 165 172   0
     21  37  73 138 152 169   0  16  47  76 130 137 154   0   3  24  30  72 104 139   0
      9  40  90 106 134 151   0  15  58  60  74 111 150 163  18  42  79 144 146 152   0
     25  38  65  99 122 160   0  17  42  75 129 170 172   0
"""

_nm = np.array(_NM_TEXT.split(), dtype=np.int32).reshape(FTX_LDPC_M, 7)

kFTX_LDPC_Nm = [[int(v) for v in row] for row in _nm]
kFTX_LDPC_Num_rows = [int((row > 0).sum()) for row in _nm]


def _derive_mn():
    mn = [[] for _ in range(FTX_LDPC_N)]
    for m in range(FTX_LDPC_M):
        for v in _nm[m]:
            if v:
                mn[int(v) - 1].append(m + 1)
    assert all(len(r) == 3 for r in mn)
    return mn


kFTX_LDPC_Mn = _derive_mn()


def parity_check_matrix() -> np.ndarray:
    """H as a dense uint8 [83,174] matrix."""
    h = np.zeros((FTX_LDPC_M, FTX_LDPC_N), dtype=np.uint8)
    for m in range(FTX_LDPC_M):
        for v in _nm[m]:
            if v:
                h[m, int(v) - 1] ^= 1
    return h


def generator_matrix() -> np.ndarray:
    """Systematic generator rows G[83,91] over GF(2): parity = G @ msg (mod 2).

    Gauss-Jordan on [B | A] with H = [A | B]; equals the 83x91 matrix of the FT8
    protocol (checked against the reference's table in tests/test_oracle_vs_reference.py).
    """
    h = parity_check_matrix()
    m = np.concatenate([h[:, FTX_LDPC_K:], h[:, :FTX_LDPC_K]], axis=1)
    for c in range(FTX_LDPC_M):
        p = c + int(np.argmax(m[c:, c]))
        if not m[p, c]:
            raise ValueError("parity part of H is singular")
        if p != c:
            m[[c, p]] = m[[p, c]]
        rows = np.nonzero(m[:, c])[0]
        rows = rows[rows != c]
        m[rows] ^= m[c]
    return m[:, FTX_LDPC_M:].copy()
