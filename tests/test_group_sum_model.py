"""CPU model of the self-validating group sums of the persistent kernel (slam3d_gx_b200/csrc/icp.cu, "Self-validating group sums"):
every CTA adds, per slot, its normalised (hi, lo) partial sums to two 64-bit words with ONE atomic each, and every atomic also
carries +1 in bits 48.. of the word.  A reader takes a word as complete when the count it decodes equals the number of CTAs of the
group.  The model replays the arithmetic on wrapping 64-bit integers, in arbitrary arrival orders, and checks the two properties
the kernel relies on: the decoded count is exactly the number of arrivals at every moment (so a word is never taken as complete
early), and the decoded value of a complete word is the exact sum, whatever the order."""
import random

import pytest

MASK = (1 << 64) - 1
ONE = 1 << 48


def to_i64(u):
    u &= MASK
    return u - (1 << 64) if u >> 63 else u


def add_wrap(word, v):          # atomicAdd on a u64 word holding a two's-complement value
    return (word + (v & MASK)) & MASK


def count_hi(word):             # ((wh + (one >> 1)) >> 48), arithmetic shift on int64
    return (to_i64(word) + (ONE >> 1)) >> 48


def count_lo(word):             # (wl >> 48): lo contributions are non-negative
    return to_i64(word) >> 48


def normalise(hi, lo):          # hi += lo >> 32; lo &= 0xffffffff   (per CTA, before the atomics)
    return hi + (lo >> 32), lo & 0xffffffff


@pytest.mark.parametrize("n_ctas", [2, 37, 148, 1024])
def test_counts_and_sums_for_any_arrival_order(n_ctas):
    rng = random.Random(1234 + n_ctas)
    for trial in range(200):
        # per-CTA raw sums as the warps hand them over: hi signed, lo a sum of many 32-bit pieces (up to 2^45)
        bound_hi = (1 << 46) // n_ctas           # |total hi| stays below 2^47: up to 2^30 terms of < 2^49 each
        raw = [(rng.randint(-bound_hi, bound_hi), rng.randint(0, (1 << 45) - 1)) for _ in range(n_ctas)]
        if trial == 0:                            # extremes
            raw = [(-bound_hi, (1 << 45) - 1)] * n_ctas
        if trial == 1:
            raw = [(bound_hi, 0)] * n_ctas
        parts = [normalise(h, l) for h, l in raw]
        assert all(0 <= l < (1 << 32) for _, l in parts)
        true_total = sum((h << 32) + l for h, l in raw)
        order = list(range(n_ctas))
        rng.shuffle(order)
        wh = wl = 0
        for arrived, c in enumerate(order, start=1):
            h, l = parts[c]
            wh = add_wrap(wh, h + ONE)
            wl = add_wrap(wl, l + ONE)
            assert count_hi(wh) == arrived and count_lo(wl) == arrived     # never complete early, never miscounted
        hi = to_i64(wh) - (n_ctas << 48)
        lo = to_i64(wl) - (n_ctas << 48)
        assert (hi << 32) + lo == true_total
        # what the solve converts: exact as long as both parts are below 2^53
        assert abs(hi) < (1 << 53) and 0 <= lo < (1 << 53)


def test_three_buffer_rotation_never_reuses_a_buffer_that_can_still_be_read():
    """Epoch e adds to buffer e % 3 while stragglers may still read buffer (e - 1) % 3; rank 0 zeroes buffer (e + 2) % 3 = (e - 1) % 3
    only after it has seen epoch e complete, i.e. after every CTA has finished reading epoch e - 1."""
    for e in range(1, 50):
        adding, reading, zeroing = e % 3, (e - 1) % 3, (e + 2) % 3
        assert adding != reading and zeroing == reading      # the buffer zeroed at the end of epoch e is the one read during epoch e
        assert (e + 2) % 3 != (e + 1) % 3                    # and it is not the one epoch e + 1 adds to
