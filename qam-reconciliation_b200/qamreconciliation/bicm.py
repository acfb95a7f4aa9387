"""Gray labelling helpers (reference: qamreconciliation/bicm.pyx)."""
import numpy as np


def generate_table_s_to_b(log_order):
    """Reflected Gray table, bit k of symbol i in column k (bicm.pyx:26-41).

    Closed form of the reference's recursion: bit k of i is set iff (i >> k) mod 4 is 1 or 2 --
    the rule the reference itself uses at noisemapper.pyx:208-215."""
    log_order = int(log_order)
    if log_order <= 0:
        raise ValueError(f"log_order ({log_order}) must be a positive integer")
    i = np.arange(1 << log_order)[:, None] >> np.arange(log_order)[None, :]
    return (((i * (i + 1)) & 3) != 0).astype(np.uint8)


def generate_error_number_table(s_to_b):
    """n_err[i, j] = Hamming distance between the labels of a_i and a_j (bicm.pyx:46-66; the
    reference indexes `s_to_b.shape[i]` there, which only works for i < 2 -- this is the intent)."""
    s = np.asarray(s_to_b, dtype=np.uint8)
    return (s[:, None, :] ^ s[None, :, :]).sum(axis=2).astype(np.int64)
