#!/usr/bin/env python
"""Convert the Sobol' generator matrices of the reference (src/core/sobolmatrices.rs: SOBOL_MATRICES_32, VDC_SOBOL_MATRICES,
VDC_SOBOL_MATRICES_INV — constant data from Joe & Kuo's direction numbers, as shipped with pbrt-v3) into the binary table the
library and the oracle embed: pbrt-rs_b200/data/sobol_tables.bin.

Layout (little endian): u32 magic 'SOB1', u32 n_dimensions (1024), u32 matrix_size (52), u32 n_vdc (25), u32 n_vdc_inv (26),
u32[3] zero padding, then u32 matrices32[n_dimensions * matrix_size], u64 vdc[n_vdc * matrix_size], u64 vdc_inv[n_vdc_inv * matrix_size].
The 64-bit matrices (float64 builds only) are not used on this path and are left out.

Run here (the reference tree is not on the GPU box): python tools/make_sobol_tables.py [/root/reference]"""
import os
import re
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def block(text, name):
    start = text.index("pub const " + name)
    body = text[text.index("= [", start) + 2:text.index("\n];", start) + 2]
    return [int(tok, 16) for tok in re.findall(r"0x[0-9a-fA-F]+", body)]


def main(ref="/root/reference"):
    text = open(os.path.join(ref, "src/core/sobolmatrices.rs")).read()
    n_dims = int(re.search(r"NUM_SOBOL_DIMENSIONS: usize = (\d+)", text).group(1))
    size = int(re.search(r"SOBOL_MATRIX_SIZE: usize = (\d+)", text).group(1))
    m32 = block(text, "SOBOL_MATRICES_32")
    vdc = block(text, "VDC_SOBOL_MATRICES:")
    inv = block(text, "VDC_SOBOL_MATRICES_INV")
    assert len(m32) == n_dims * size, (len(m32), n_dims, size)
    assert len(vdc) % size == 0 and len(inv) % size == 0
    n_vdc, n_inv = len(vdc) // size, len(inv) // size
    out = os.path.join(ROOT, "pbrt-rs_b200", "data", "sobol_tables.bin")
    with open(out, "wb") as f:
        f.write(struct.pack("<8I", 0x31424F53, n_dims, size, n_vdc, n_inv, 0, 0, 0))
        f.write(struct.pack(f"<{len(m32)}I", *m32))
        f.write(struct.pack(f"<{len(vdc)}Q", *vdc))
        f.write(struct.pack(f"<{len(inv)}Q", *inv))
    print(f"{out}: {n_dims} dimensions x {size}, {n_vdc} + {n_inv} van der Corput matrices, {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    main(*sys.argv[1:2])
