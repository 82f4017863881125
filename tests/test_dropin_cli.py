"""The drop-in boundary: the reference's UNMODIFIED myyuv_cli (object compiled from /root/reference/myyuv_cli/main.cpp
against the reference's own headers, oracle/_ref/myyuv_cli_main.o) linked with this repo's libmyyuv_lib.so."""
import hashlib
import os
import pathlib
import re
import subprocess

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent
LIB = ROOT / "yuv-manipulations-2_b200" / "lib"


def class_api_symbols(path):
    out = subprocess.run(f"nm -D --defined-only {path} | c++filt", shell=True, capture_output=True, text=True).stdout
    return sorted(m.group(1) for m in re.finditer(r"^[0-9a-f]+ [TBDVWu] (myyuv::.*)$", out, re.M))


def test_exports_the_reference_class_api():
    """Same exported myyuv:: symbols (constructors, methods, the seven static registries) as the reference library."""
    mine = class_api_symbols(LIB / "libmyyuv_lib.so")
    assert len(mine) >= 60
    for must in ["myyuv::YUV::compress_map", "myyuv::YUV::decompress_map", "myyuv::YUV::bmp_to_yuv_map",
                 "myyuv::YUV::compress(unsigned short, void const*, unsigned int) const", "myyuv::YUV::decompress() const",
                 "myyuv::YUV::YUV(myyuv::BMP const&, unsigned int)", "myyuv::BMP::colorData() const"]:
        assert must in mine, must
    ref = ROOT / "oracle" / "_ref" / "serial" / "libmyyuv_lib.so"
    if ref.exists():
        assert mine == class_api_symbols(ref)


def test_unmodified_cli_links():
    if not (ROOT / "oracle" / "_ref" / "myyuv_cli_main.o").exists():
        pytest.skip("reference CLI object not built (oracle/_ref)")
    assert (LIB / "myyuv_cli").exists()
    out = subprocess.run(["ldd", str(LIB / "myyuv_cli")], capture_output=True, text=True).stdout
    assert str(LIB / "libmyyuv_lib.so") in out and "libmyyuvb200.so" in out


@pytest.mark.gpu
def test_unmodified_cli_reproduces_golden_files(golden_dir, tmp_path):
    """myyuv_cli x.bmp -to_yuv IYUV / -compress DCT 50 / -compress DCT 90 / -decompress run the sm_100a kernels and
    reproduce the reference's shipped files byte for byte (SURVEY 7.2 minimum slice)."""
    cli = LIB / "myyuv_cli"
    if not cli.exists():
        pytest.skip("drop-in CLI not built")

    def run(*args):
        r = subprocess.run([str(cli), *map(str, args)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "Success!" in r.stdout, r.stdout + r.stderr
        return r.stdout

    a, b, c, d, e = (tmp_path / n for n in ("a.myyuv", "b.myyuv", "c.myyuv", "d.myyuv", "e.myyuv"))
    run(golden_dir / "chef-with-trumpet.bmp", "-to_yuv", "IYUV", "-o", a)
    assert a.read_bytes() == (golden_dir / "chef-with-trumpet.myyuv").read_bytes()
    run(a, "-compress", "DCT", "50", "-o", b)
    assert b.read_bytes() == (golden_dir / "chef-with-trumpet-DCT-50.myyuv").read_bytes()
    run(a, "-compress", "DCT", "90", "-o", c)
    assert c.read_bytes() == (golden_dir / "chef-with-trumpet-DCT-90.myyuv").read_bytes()
    run(golden_dir / "chef-with-trumpet-DCT-50.myyuv", "-decompress", "-o", d)
    assert hashlib.sha256(d.read_bytes()).hexdigest().startswith("a95127da47152")
    run(golden_dir / "chef-with-trumpet-big-DCT-50.myyuv", "-decompress", "-o", e)
    assert hashlib.sha256(e.read_bytes()).hexdigest().startswith("5e77691911882")
    info = run(b, "-info")
    assert "Compression: 1" in info and "Width: 992" in info and "Valid: 1" in info


@pytest.mark.gpu
def test_cli_and_reference_cli_agree_on_synthetic(synth, tmp_path):
    ref_cli = ROOT / "oracle" / "_ref" / "serial" / "myyuv_cli"
    cli = LIB / "myyuv_cli"
    if not (ref_cli.exists() and cli.exists()):
        pytest.skip("CLIs not built")
    import importlib

    pkg = importlib.import_module("yuv-manipulations-2_b200")
    w, h = 640, 368 - 368 % 16
    y = pkg.YUV()
    y.header.fourcc_format = pkg.YUV.FourccFormats.IYUV
    y.header.width, y.header.height, y.header.data_size, y.header.data_pos = w, h, w * h * 3 // 2, 64
    y.data = synth.iyuv_frames_numpy(w, h, 1, 2)[0]
    src = tmp_path / "src.myyuv"
    y.dump(str(src))
    for q in (["35"], ["80", "20", "60"]):
        outs = []
        for exe, tag in ((cli, "ours"), (ref_cli, "ref")):
            o = tmp_path / f"{tag}.myyuv"
            r = subprocess.run([str(exe), str(src), "-compress", "DCT", *q, "-o", str(o)], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            outs.append(o.read_bytes())
        assert outs[0] == outs[1]
        # cross decode: the reference CLI decodes our file, our CLI decodes the reference's
        d1, d2 = tmp_path / "d1.myyuv", tmp_path / "d2.myyuv"
        subprocess.run([str(ref_cli), str(tmp_path / "ours.myyuv"), "-decompress", "-o", str(d1)], check=True, capture_output=True, timeout=300)
        subprocess.run([str(cli), str(tmp_path / "ref.myyuv"), "-decompress", "-o", str(d2)], check=True, capture_output=True, timeout=300)
        assert d1.read_bytes() == d2.read_bytes()


@pytest.mark.gpu
def test_cli_converts_24bit_bmp_like_the_reference_cli(synth, tmp_path):
    """A 24-bit BMP through the unmodified CLI linked to the replacement library and through the reference's own CLI
    (Release build: the 32-bit assert is compiled out): identical .myyuv files."""
    import struct

    import numpy as np

    ref_cli = ROOT / "oracle" / "_ref" / "serial" / "myyuv_cli"
    cli = LIB / "myyuv_cli"
    if not (ref_cli.exists() and cli.exists()):
        pytest.skip("CLIs not built")
    # bottom-up, top-down, and the reversed pixel order colorData() produces for a negative width (myyuv_bmp.cpp:89-94)
    for w_signed, h_signed in ((64, 48), (40, -24), (-40, 24)):
        w, h = abs(w_signed), abs(h_signed)
        px = np.ascontiguousarray(synth.bgrx_frames_numpy(w, h, 1, 4)[0].reshape(-1, 4)[:, :3]).tobytes()
        header = b"BM" + struct.pack("<IHHIIiiHHIIiiII", 54 + len(px), 0, 0, 54, 40, w_signed, h_signed, 1, 24, 0, 0, 2835, 2835, 0, 0)
        assert len(header) == 54
        bmp = tmp_path / f"in{w_signed}_{h_signed}.bmp"
        bmp.write_bytes(header + px)
        outs = []
        for exe, tag in ((cli, "ours"), (ref_cli, "ref")):
            o = tmp_path / f"{tag}{w_signed}_{h_signed}.myyuv"
            r = subprocess.run([str(exe), str(bmp), "-to_yuv", "IYUV", "-o", str(o)], capture_output=True, text=True, timeout=300)
            assert r.returncode == 0 and "Success!" in r.stdout, r.stdout + r.stderr
            outs.append(o.read_bytes())
        assert outs[0] == outs[1] and len(outs[0]) == 64 + w * h * 3 // 2
        # the Python mirror of the class API takes the same orientations (it used to refuse width < 0, height > 0)
        import importlib

        pkg = importlib.import_module("yuv-manipulations-2_b200")
        py = tmp_path / f"py{w_signed}_{h_signed}.myyuv"
        pkg.YUV(pkg.BMP(str(bmp)), pkg.YUV.FourccFormats.IYUV).dump(str(py))
        assert py.read_bytes() == outs[1]


@pytest.mark.gpu
def test_registry_plugin_over_the_unmodified_reference_library(tmp_path):
    """INTEGRATION.md form B: the reference's own CLI and library, with the three hot-path registry slots overridden by the
    LD_PRELOADed plugin (csrc/registry_plugin.cpp).  Same process image, slots switched by MYYUVB_PLUGIN: the outputs of
    the CPU reference and of the sm_100a kernels must be the same files, and both equal the shipped golden files."""
    ref_cli = ROOT / "oracle" / "_ref" / "serial" / "myyuv_cli"
    plugin = ROOT / "oracle" / "_ref" / "plugin" / "libmyyuvb200_plugin.so"
    golden = ROOT / "oracle" / "_ref" / "golden"
    if not (ref_cli.exists() and plugin.exists() and (golden / "chef-with-trumpet.bmp").exists()):
        pytest.skip("reference CLI / plugin / golden images not built (make -C oracle ref plugin)")

    def run(tag, on, *args):
        env = dict(os.environ, LD_PRELOAD=str(plugin), MYYUVB_PLUGIN="1" if on else "0", MYYUVB_PLUGIN_VERBOSE="1")
        r = subprocess.run([str(ref_cli), *[str(a) for a in args]], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "Success!" in r.stdout, r.stdout + r.stderr
        assert ("registry slots IYUV / DCT overridden" in r.stderr) == on
        return r

    outs = {}
    for on in (False, True):
        t = "gpu" if on else "cpu"
        a, b, c, d = (tmp_path / f"{t}_{n}.myyuv" for n in "abcd")
        run(t, on, golden / "chef-with-trumpet.bmp", "-to_yuv", "IYUV", "-o", a)
        run(t, on, a, "-compress", "DCT", "50", "-o", b)
        run(t, on, a, "-compress", "DCT", "90", "20", "75", "-o", c)
        run(t, on, golden / "chef-with-trumpet-big-DCT-50.myyuv", "-decompress", "-o", d)
        outs[on] = [p.read_bytes() for p in (a, b, c, d)]
    assert outs[True] == outs[False]
    assert outs[True][0] == (golden / "chef-with-trumpet.myyuv").read_bytes()
    assert outs[True][1] == (golden / "chef-with-trumpet-DCT-50.myyuv").read_bytes()
    assert hashlib.sha256(outs[True][3]).hexdigest().startswith("5e77691911882")
