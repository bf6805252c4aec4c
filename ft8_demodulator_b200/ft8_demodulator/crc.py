"""CRC-14 helpers with the reference's names (`ft8_tools/ft8_demodulator/crc.py:11-79`).

These byte-level utilities run on the host for API compatibility; the decode path itself checks the
CRC inside the CUDA decode kernel (csrc/decode.cu) and never calls them.
"""
CRC_WIDTH = 14
CRC_POLYNOMIAL = 0x2757
TOPBIT = 1 << (CRC_WIDTH - 1)


def compute_crc(message, num_bits: int) -> int:
    """CRC-14 over the first `num_bits` bits of `message`, MSB first (crc.py:11-39)."""
    rem = 0
    for i in range(num_bits):
        if i % 8 == 0:
            rem ^= message[i // 8] << (CRC_WIDTH - 8)
        rem = ((rem << 1) ^ CRC_POLYNOMIAL) if (rem & TOPBIT) else (rem << 1)
        rem &= (TOPBIT << 1) - 1
    return rem


def extract_crc(a91) -> int:
    """The 14 CRC bits stored after the 77 payload bits (crc.py:41-54)."""
    return ((a91[9] & 0x07) << 11) | (a91[10] << 3) | (a91[11] >> 5)


def add_crc(payload, a91) -> None:
    """Fill a91 (12 bytes) with payload + CRC (crc.py:56-79)."""
    for i in range(10):
        a91[i] = payload[i]
    a91[9] &= 0xF8
    a91[10] = 0
    a91[11] = 0
    c = compute_crc(a91, 82)
    a91[9] |= (c >> 11)
    a91[10] = (c >> 3) & 0xFF
    a91[11] = (c << 5) & 0xE0
