"""Unit-level pins of the oracle against the reference's own known answers and
constant tables.  CPU only."""
import ctypes
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VEC = bytes([0xB1, 0xED, 0x3B, 0xC1])   # the reference's bit-reader test vector (bitstream.c:4545)


def test_unsigned_reads_known_answers(oracle):
    # reference src/bitstream.c:4864-4868: read(2)=2, read(3)=6, read(5)=7, read(3)=5, read(19)=0x53BC1
    L = oracle.lib()
    pos, got = 0, []
    for n in (2, 3, 5, 3, 19):
        got.append(L.dvda_oracle_read_bits(VEC, len(VEC), pos, n))
        pos += n
    assert got == [2, 6, 7, 5, 0x53BC1]


def test_signed_reads_known_answers(oracle):
    # reference src/bitstream.c:4940-4944: -2, -2, 7, -3, -181311
    L = oracle.lib()
    pos, got = 0, []
    for n in (2, 3, 5, 3, 19):
        got.append(L.dvda_oracle_read_signed(VEC, len(VEC), pos, n))
        pos += n
    assert got == [-2, -2, 7, -3, -181311]


# code strings -> value, as listed in SURVEY.md A.14 from src/mlp_codebook{1,2,3}.json
def _codebook(n):
    low = {"0" * (8 - v) + "1": v for v in range(0, 7)}
    if n == 1:
        mid = {"100": 7, "101": 8, "110": 9, "111": 10}
        hi0 = 11
    elif n == 2:
        mid = {"10": 7, "11": 8}
        hi0 = 9
    else:
        mid = {"1": 7}
        hi0 = 8
    high = {"01" + "0" * k + "1": hi0 + k for k in range(0, 7)}
    book = dict(low)
    book.update(mid)
    book.update(high)
    book["010000000"] = -1
    book["000000000"] = -1
    return book


def _pack(bits):
    bits = bits + "0" * (-len(bits) % 8) + "0" * 16
    return bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8))


@pytest.mark.parametrize("cb", [1, 2, 3])
def test_huffman_codebooks(oracle, cb):
    L = oracle.lib()
    book = _codebook(cb)
    assert len(book) == {1: 20, 2: 18, 3: 17}[cb]          # code counts of the three JSON files
    ref_json = os.path.join("/root/reference/src", "mlp_codebook%d.json" % cb)
    if os.path.exists(ref_json):                             # in the build container: check against the file itself
        j = json.load(open(ref_json))
        from_file = {"".join(map(str, j[i])): j[i + 1] for i in range(0, len(j), 2)}
        assert from_file == book
    for code, value in book.items():
        for lead in ("", "1", "01"):                         # at several bit offsets
            buf = _pack(lead + code + "1")
            n = ctypes.c_uint()
            got = L.dvda_oracle_huffman(buf, len(buf), len(lead), cb, ctypes.byref(n))
            assert got == value, (cb, code)
            if value >= 0:
                assert n.value == len(code)


def test_crc8_table_is_poly_0x63(oracle):
    L = oracle.lib()
    # spot values of the table at reference src/mlp.c:1363-1395
    assert [L.dvda_oracle_crc8_table(i) for i in (0, 1, 2, 3, 4, 0x80, 0xFF)] == [0x00, 0x63, 0xC6, 0xA5, 0xEF, 0xC8, 0x70]


def test_pcm_permutation_is_a_permutation(oracle):
    L = oracle.lib()
    for bits in (16, 24):
        for ch in range(1, 7):
            n = bits // 8 * ch * 2
            tab = ctypes.create_string_buffer(36)
            L.dvda_oracle_pcm_permutation(bits, ch, tab)
            assert sorted(tab.raw[:n]) == list(range(n))
    # 16-bit stereo is plain big-endian (reference src/pcm.c:105)
    tab = ctypes.create_string_buffer(36)
    L.dvda_oracle_pcm_permutation(16, 2, tab)
    assert list(tab.raw[:8]) == [1, 0, 3, 2, 5, 4, 7, 6]
    # 24-bit stereo: four (high, middle) pairs, then the four low bytes (src/pcm.c:121-122)
    L.dvda_oracle_pcm_permutation(24, 2, tab)
    assert list(tab.raw[:12]) == [2, 1, 5, 4, 8, 7, 11, 10, 0, 3, 6, 9]
