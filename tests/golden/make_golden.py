"""Extract the reference's committed golden vectors into small .npz fixtures.

Source: /root/reference/examples/*.h5 (HDF5 superblock v0, contiguous little-endian datasets at
fixed byte offsets, SURVEY.md Appendix B -- h5py is not available). Run in the build container
only; the GPU box never sees /root/reference, it reads tests/golden/*.npz.

    python tests/golden/make_golden.py
"""

import pathlib

import numpy as np

REF = pathlib.Path("/root/reference/examples")
OUT = pathlib.Path(__file__).resolve().parent

LAYOUT = {
    "Line1d_Cuspy_Laplace": 1000,
    "Line1d_Cuspy_Laplace_Nopassing": 1000,
    "Line1d_Cuspy_Quartic": 1000,
    "Line1d_SemiSmooth_Laplace": 1000,
    "Line2d_Cuspy_Laplace": 1000,
    "Line1d_Cuspy_Laplace_LongRange": 200,
    "Particles_Cuspy": 2000,  # first dataset is named u_frame in this file (same layout)
}


def read(name: str, n: int):
    buf = (REF / f"{name}.h5").read_bytes()
    assert buf[:8] == b"\x89HDF\r\n\x1a\n"
    x_frame = np.frombuffer(buf, "<f8", n, 2048)
    f_frame = np.frombuffer(buf, "<f8", n, 2048 + 8 * n)
    S = np.frombuffer(buf, "<i8", n, 2048 + 16 * n)
    return x_frame, f_frame, S


def read_thermal():
    """examples/Line1d_System_Cuspy_Laplace_RandomForcing.h5: x_frame, f_frame, t_insta (500)."""
    buf = (REF / "Line1d_System_Cuspy_Laplace_RandomForcing.h5").read_bytes()
    assert buf[:8] == b"\x89HDF\r\n\x1a\n"
    return (np.frombuffer(buf, "<f8", 500, 2048), np.frombuffer(buf, "<f8", 500, 6048),
            np.frombuffer(buf, "<f8", 500, 10048))


if __name__ == "__main__":
    x_frame, f_frame, t_insta = read_thermal()
    assert np.allclose(np.diff(x_frame), 5.0) and np.all(t_insta > 0)
    np.savez_compressed(OUT / "Line1d_System_Cuspy_Laplace_RandomForcing.npz", x_frame=x_frame,
                        f_frame=f_frame, t_insta=t_insta)
    print("Line1d_System_Cuspy_Laplace_RandomForcing", 500, "t_insta[-1] =", t_insta[-1])
    for name, n in LAYOUT.items():
        x_frame, f_frame, S = read(name, n)
        assert np.all(np.diff(x_frame) >= 0) and np.all(np.isfinite(f_frame))
        np.savez_compressed(OUT / f"{name}.npz", x_frame=x_frame, f_frame=f_frame, S=S)
        print(name, n, "sum(S) =", int(S.sum()), "x_frame[-1] =", x_frame[-1])
