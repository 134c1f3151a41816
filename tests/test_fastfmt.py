"""Output formatter: fmt_f6 of minimod_b200/host/fastfmt.h must print exactly what the reference's
fprintf("%f") prints (src/mod.c:685,703) -- exhaustively for small denominators, plus random 32-bit counts and
dyadic values where exact round-half-to-even ties occur; and the writer built on it must still reproduce goldens."""
import ctypes as C


def test_fmt_f6_equals_printf(host_lib):
    msg = C.create_string_buffer(256)
    bad = host_lib.mmh_fastfmt_selftest(1024, 2_000_000, msg, 256)
    assert bad == 0, msg.value.decode()


def test_freq_writer_matches_oracle_binary_text():
    """The writer is exercised byte-for-byte by every golden test; here: a few exact-tie values through the writer."""
    import os, tempfile
    import numpy as np
    from minimod_b200 import _native as N
    host = N.load_host()
    rows = np.zeros(6, dtype=N.FREQ_DTYPE)
    vals = [(1, 128), (3, 128), (5, 1024), (1, 3), (7, 7), (0, 5)]          # 1/128 = 0.0078125: a tie at 6 decimals
    for i, (m, c) in enumerate(vals):
        rows[i] = (0, 100 + i, c, m, 0, -1, i & 1, 0, 0)
    names = (C.c_char_p * 1)(b"chrT")
    codes = (C.c_char_p * 256)(*([b"m"] * 256))
    for bed in (0, 1):
        with tempfile.NamedTemporaryFile(suffix=".tsv", delete=False) as tf:
            path = tf.name
        try:
            recs = rows.ctypes.data_as(C.POINTER(N.MmcFreqRec))
            assert host.mmh_write_freq(path.encode(), bed, 0, 0, 1, names, recs, len(rows), 256, codes) == 0
            text = open(path).read().splitlines()
        finally:
            os.unlink(path)
        body = text[0 if bed else 1:]
        for line, (m, c) in zip(body, vals):
            f = line.split("\t")
            want = "%f" % ((m * 100 / c) if bed else (m / c))
            assert (f[10] if bed else f[6]) == want, (line, want)
