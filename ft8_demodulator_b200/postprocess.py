"""Opt-in result post-processing the reference does not have: de-duplication of repeated
decodes and 77-bit -> text unpacking.  Host-side, outside the hot path, and NOT part of the drop-in surface: the
default output of decode_ft8_message / decode_slots stays exactly the reference's list.

Parity: unpinned - the reference stops at the 10-byte payload (ft8_decode.py:383-391), so there is nothing in it to
compare against.  The unpacking follows the published FT8 message layout (Franke/Somerville/Taylor, "The FT4 and FT8
Communication Protocols", QEX 2020, section 4 / table 1); tests/test_host_logic.py checks it by pack -> unpack round
trips over the message types handled here.
"""
from __future__ import annotations

NTOKENS = 2063592
MAX22 = 4194304
MAXGRID4 = 32400

_A1 = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_A2 = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_A3 = "0123456789"
_A4 = " ABCDEFGHIJKLMNOPQRSTUVWXYZ"
_FREE = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ+-./?"
_C58 = " 0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ/"


def payload_to_int77(payload10) -> int:
    """First 77 bits of the 10-byte payload, MSB first (the last 3 bits of byte 9 are padding, ft8_decode.py:271)."""
    b = bytes(payload10)
    if len(b) != 10:
        raise ValueError("payload must be 10 bytes")
    return int.from_bytes(b, "big") >> 3


def int77_to_payload(v: int) -> bytes:
    return ((v & ((1 << 77) - 1)) << 3).to_bytes(10, "big")


def _bits(v: int, start: int, length: int) -> int:
    """`length` bits of the 77-bit word starting `start` bits from its most significant end."""
    return (v >> (77 - start - length)) & ((1 << length) - 1)


# ------------------------------------------------------------------------------------------ 28-bit call field
def unpack_c28(n28: int) -> str:
    if n28 < NTOKENS:
        if n28 == 0:
            return "DE"
        if n28 == 1:
            return "QRZ"
        if n28 == 2:
            return "CQ"
        if n28 <= 1002:
            return "CQ %03d" % (n28 - 3)
        if n28 <= 532443:
            n = n28 - 1003
            s = ""
            for _ in range(4):
                s = _A4[n % 27] + s
                n //= 27
            return "CQ " + s.strip()
        return "<token %d>" % n28
    n = n28 - NTOKENS
    if n < MAX22:
        return "<...>"                       # 22-bit hash of a call sign announced earlier; no hash table kept here
    n -= MAX22
    c6 = _A4[n % 27]; n //= 27
    c5 = _A4[n % 27]; n //= 27
    c4 = _A4[n % 27]; n //= 27
    c3 = _A3[n % 10]; n //= 10
    c2 = _A2[n % 36]; n //= 36
    if n >= 37:
        return "<bad call>"
    return (_A1[n] + c2 + c3 + c4 + c5 + c6).strip()


def pack_c28(call: str) -> int:
    call = call.strip().upper()
    if call == "DE":
        return 0
    if call == "QRZ":
        return 1
    if call == "CQ":
        return 2
    if call.startswith("CQ "):
        rest = call[3:].strip()
        if rest.isdigit() and len(rest) == 3:
            return 3 + int(rest)
        if 1 <= len(rest) <= 4 and rest.isalpha():
            n = 0
            for ch in rest.rjust(4):
                n = n * 27 + _A4.index(ch)
            return 1003 + n
        raise ValueError("cannot pack %r" % call)
    # standard call sign: align so that the digit is the third character
    c = call
    if len(c) >= 2 and c[1].isdigit() and not (len(c) >= 3 and c[2].isdigit()):
        c = " " + c
    c = c.ljust(6)
    if len(c) != 6:
        raise ValueError("not a standard call sign: %r" % call)
    try:
        n = _A1.index(c[0])
        n = n * 36 + _A2.index(c[1])
        n = n * 10 + _A3.index(c[2])
        n = n * 27 + _A4.index(c[3])
        n = n * 27 + _A4.index(c[4])
        n = n * 27 + _A4.index(c[5])
    except ValueError:
        raise ValueError("not a standard call sign: %r" % call) from None
    return NTOKENS + MAX22 + n


# ------------------------------------------------------------------------------------------ 15-bit grid / report field
def unpack_g15(g15: int, ir: int) -> str:
    if g15 <= MAXGRID4:
        n = g15
        d4 = n % 10; n //= 10
        d3 = n % 10; n //= 10
        l2 = n % 18; n //= 18
        if n >= 18:
            return "<bad grid>"
        grid = chr(65 + n) + chr(65 + l2) + str(d3) + str(d4)
        return ("R " + grid) if ir else grid
    irpt = g15 - MAXGRID4
    if irpt == 1:
        return ""
    if irpt == 2:
        return "RRR"
    if irpt == 3:
        return "RR73"
    if irpt == 4:
        return "73"
    return ("R" if ir else "") + "%+03d" % (irpt - 35)


def pack_g15(field: str):
    """-> (g15, ir)"""
    f = field.strip().upper()
    if f == "":
        return MAXGRID4 + 1, 0
    if f == "RRR":
        return MAXGRID4 + 2, 0
    if f == "RR73":
        return MAXGRID4 + 3, 0
    if f == "73":
        return MAXGRID4 + 4, 0
    ir = 0
    if f.startswith("R ") and len(f) == 6:
        ir, f = 1, f[2:]
    if len(f) == 4 and f[0] in "ABCDEFGHIJKLMNOPQR" and f[1] in "ABCDEFGHIJKLMNOPQR" and f[2:].isdigit():
        return ((ord(f[0]) - 65) * 18 + (ord(f[1]) - 65)) * 100 + int(f[2:]), ir
    if f.startswith("R") and len(f) > 1 and f[1] in "+-":
        ir, f = 1, f[1:]
    if f[0] in "+-" and f[1:].isdigit():
        rpt = int(f)
        if not -30 <= rpt <= 99:
            raise ValueError("report out of range: %r" % field)
        return MAXGRID4 + 35 + rpt, ir
    raise ValueError("cannot pack %r" % field)


# ------------------------------------------------------------------------------------------ whole messages
def pack77_standard(call_to: str, call_de: str, extra: str = "", i3: int = 1) -> bytes:
    """Type 1 (or 2) message -> 10-byte payload.  For tests of unpack77; not a general FT8 encoder."""
    def split(call):
        suffix = "/R" if i3 == 1 else "/P"
        if call.upper().endswith(suffix):
            return call[:-2], 1
        return call, 0
    a, ipa = split(call_to)
    b, ipb = split(call_de)
    g15, ir = pack_g15(extra)
    v = pack_c28(a)
    v = (v << 1) | ipa
    v = (v << 28) | pack_c28(b)
    v = (v << 1) | ipb
    v = (v << 1) | ir
    v = (v << 15) | g15
    v = (v << 3) | i3
    return int77_to_payload(v)


def pack77_free_text(text: str) -> bytes:
    t = text.upper()[:13].rjust(13)
    n = 0
    for ch in t:
        n = n * 42 + _FREE.index(ch)
    return int77_to_payload(n << 6)            # n3 = 0, i3 = 0


def unpack77(payload10) -> str:
    """Human-readable text of a decoded payload; message types this does not cover come back as 'i3.n3:<hex>'."""
    v = payload_to_int77(payload10)
    i3 = _bits(v, 74, 3)
    if i3 == 0:
        n3 = _bits(v, 71, 3)
        if n3 == 0:
            n = _bits(v, 0, 71)
            s = ""
            for _ in range(13):
                s = _FREE[n % 42] + s
                n //= 42
            return s.strip()
        if n3 == 5:
            return "%018X" % _bits(v, 0, 71)
        return "0.%d:%s" % (n3, bytes(payload10).hex())
    if i3 in (1, 2):
        suffix = "/R" if i3 == 1 else "/P"
        a = unpack_c28(_bits(v, 0, 28)) + (suffix if _bits(v, 28, 1) else "")
        b = unpack_c28(_bits(v, 29, 28)) + (suffix if _bits(v, 57, 1) else "")
        g = unpack_g15(_bits(v, 59, 15), _bits(v, 58, 1))
        return " ".join(x for x in (a, b, g) if x)
    if i3 == 4:
        n = _bits(v, 12, 58)
        call = ""
        for _ in range(11):
            call = _C58[n % 38] + call
            n //= 38
        call = call.strip()
        if _bits(v, 73, 1):
            return "CQ " + call
        first, second = ("<...>", call) if _bits(v, 70, 1) == 0 else (call, "<...>")
        tail = ["", "RRR", "RR73", "73"][_bits(v, 71, 2)]
        return " ".join(x for x in (first, second, tail) if x)
    return "%d:%s" % (i3, bytes(payload10).hex())


def dedupe(results: list) -> list:
    """Drop repeated payloads from a decode_ft8_message result list, keeping the first (= highest sync score)
    occurrence of each; the same burst is usually found at several neighbouring (time, frequency) offsets."""
    seen = set()
    out = []
    for r in results:
        key = bytes(r[0].payload)
        if key not in seen:
            seen.add(key)
            out.append(r)
    return out


def annotate(results: list, unique: bool = True) -> list:
    """[(text, payload hex, time_sec, freq_hz, score)] for a decode_ft8_message result list."""
    return [(unpack77(m.payload), bytes(m.payload).hex(), t, f, s) for m, _st, t, f, s in (dedupe(results) if unique else results)]
