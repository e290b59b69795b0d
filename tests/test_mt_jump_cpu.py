"""Host-side algebra of the MT19937 jump-ahead (mpb_rng.cu): the characteristic polynomial the library derives by
Berlekamp-Massey and its x^J mod phi are checked against the defining property on a NumPy-generated word stream:
x[m+J] = XOR over the set bits i of x[m+i].  No GPU needed (mpb_mt19937_jump_poly is host-only)."""
import numpy as np

from magphase_b200 import _lib


def _stream(seed, n):
    """x[0..n): the 624 state words NumPy holds after seeding, continued by the twister's recurrence."""
    x = np.zeros(n, dtype=np.uint32)
    x[:624] = np.random.RandomState(seed).get_state()[1]
    k = 624
    while k < n:
        m = min(227, n - k)
        a, b = x[k - 624:k - 624 + m], x[k - 623:k - 623 + m]
        y = (a & np.uint32(0x80000000)) | (b & np.uint32(0x7fffffff))
        x[k:k + m] = x[k - 227:k - 227 + m] ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), np.uint32(0x9908b0df), np.uint32(0))
        k += m
    return x


def _poly_bits(n_words):
    out = np.zeros(624, dtype=np.uint32)
    _lib.check(_lib.lib().mpb_mt19937_jump_poly(int(n_words), _lib.ptr(out)))
    return np.flatnonzero(np.unpackbits(out.view(np.uint8), bitorder='little'))


def test_characteristic_polynomial_annihilates_the_stream():
    phi = np.r_[_poly_bits(0), 19937]                       # leading term added back
    assert phi[0] == 0 and phi.size > 100
    x = _stream(1, 19937 + 2000)
    for m in (1, 2, 623, 1999):
        assert np.bitwise_xor.reduce(x[m + phi]) == 0


def test_jump_polynomial_reaches_the_segment_start():
    J = 256 * 624                                          # MT_SEG_WORDS
    x = _stream(12345, J + 19937 + 1400)
    g = _poly_bits(J)
    assert g.max() < 19937
    for m in (1, 5, 624, 1300):
        assert np.bitwise_xor.reduce(x[m + g]) == x[m + J]
    g3 = _poly_bits(3)                                     # tiny jumps are monomials
    assert g3.tolist() == [3]
