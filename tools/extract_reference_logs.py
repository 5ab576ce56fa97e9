#!/usr/bin/env python3
"""Transcribe the known answers the reference ships in its own example logs into tests/golden/reference_logs.json.

The reference (Fortran + MPI + HDF5 + CFITSIO) cannot be built in this image, so these logged numbers — set-up scalars
and whole-run averages printed by LaRT itself — are the golden vectors of this repository.  Run where /root/reference
exists:  python tools/extract_reference_logs.py [/root/reference]"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
num = r"([-+]?\d+\.?\d*[Ee]?[-+]?\d*)"


def grab(path, first, last, pattern, cast=float):
    """first match of `pattern` in lines first..last (1-based) of examples/<path>; returns (value, 'path:line')"""
    with open(os.path.join(ref, "examples", path)) as fh:
        lines = fh.read().splitlines()
    for i in range(first - 1, min(last, len(lines))):
        m = re.search(pattern, lines[i])
        if m:
            return {"value": cast(m.group(1)), "source": "examples/%s:%d" % (path, i + 1), "text": lines[i].strip()}
    raise SystemExit("pattern %r not found in %s:%d-%d" % (pattern, path, first, last))


def pdf_text(path):
    """Text of a (LaTeX-made) PDF under docs/: inflate the content streams and join the TJ/Tj string operands."""
    import zlib
    data = open(os.path.join(ref, "docs", path), "rb").read()
    pages = []
    for m in re.finditer(rb"stream\r?\n", data):
        raw = data[m.end():data.find(b"endstream", m.end())]
        try:
            txt = zlib.decompress(raw)
        except Exception:
            continue
        if b"TJ" not in txt and b"Tj" not in txt:
            continue
        out = []
        for mm in re.finditer(rb"\[(.*?)\]\s*TJ|\((.*?)\)\s*Tj", txt, re.S):
            if mm.group(1) is not None:
                for a, b in re.findall(rb"\(((?:\\.|[^\\)])*)\)|(-?\d+\.?\d*)", mm.group(1)):
                    out.append(b" " if (b and float(b) < -200) else a)
            else:
                out.append(mm.group(2))
        pages.append(b"".join(out).decode("latin1"))
    return " ".join(pages)


def grab_doc(text, pattern, source, cast=float):
    m = re.search(pattern, text)
    if not m:
        raise SystemExit("pattern %r not found in %s" % (pattern, source))
    return {"value": cast(m.group(1).replace(" ", "").replace(":", ".")), "source": source, "text": m.group(0).strip()}


_doc = pdf_text("LaRT_AMR_description.pdf")
_d = r"(\d+:\d+)"  # a decimal number as the PDF holds it: '7:402'
_src = "docs/LaRT_AMR_description.pdf section 11 (validation table: 101^3 Cartesian sphere, T = 1e4 K, central point source, 1e6 photons)"
_row2 = r"102\d+:\d+\\00210-3" + _d + r"\\00210-30\.970\\006" + _d + "kms"   # tau0 = 1e2: J_AMR, J_CAR, ratio, peak velocity
_row4 = r"104\d+:\d+\\00210-3" + _d + r"\\00210-30\.965\\006" + _d + "kms"   # tau0 = 1e4


def grab_doc2(text, pattern, group, source, cast=float):
    m = re.search(pattern, text)
    if not m:
        raise SystemExit("pattern %r not found in %s" % (pattern, source))
    return {"value": cast(m.group(group).replace(":", ".")), "source": source, "text": m.group(0)}


out = {
    # Known answers of the Cartesian runs the reference documents beside its AMR runs: peak of the normalised J_out(x) and
    # the velocity of the peak bin (the default 121-bin grid on [-9, 9])
    "doc_car_sphere_101": {
        "Jout_max_tau1e2": grab_doc2(_doc, _row2, 1, _src, lambda t: float(t) * 1e-3),
        "Jout_max_tau1e4": grab_doc2(_doc, _row4, 1, _src, lambda t: float(t) * 1e-3),
        "peak_kms_tau1e2": grab_doc2(_doc, _row2, 2, _src),
        "peak_kms_tau1e4": grab_doc2(_doc, _row4, 2, _src),
    },
    "amr_sphere_generic_amr_1M": {  # log_amr_1M.txt: the octree twin (178 480 leaves) of the 64^3 sphere below
        "nleaf": grab("amr_sphere_generic/log_amr_1M.txt", 1, 40, r"AMR nleaf\s*:\s*" + num, int),
        "voigt_a": grab("amr_sphere_generic/log_amr_1M.txt", 1, 40, r"AMR voigt_a\s*:\s*" + num),
        "N_HI_pole": grab("amr_sphere_generic/log_amr_1M.txt", 1, 40, r"AMR N\(HI\)_pole\s*:\s*" + num),
        "tau_pole": grab("amr_sphere_generic/log_amr_1M.txt", 1, 40, r"AMR tau_pole\s*:\s*" + num),
        "nphotons": grab("amr_sphere_generic/log_amr_1M.txt", 1, 40, r"Total number of photons\s*:\s*" + num),
        "mean_nscatt": grab("amr_sphere_generic/log_amr_1M.txt", 1, 60, r"Average Number of scattering\s*:\s*" + num),
        "wall_minutes": grab("amr_sphere_generic/log_amr_1M.txt", 1, 60, r"Total Excution Time\s*:\s*" + num),
    },
    "sphere_peel_t1tau3": {  # examples/sphere_peel/out.txt, input t1tau3.in: T = 10 K, tau0 = 1e3, 201^3, 1e7 photons
        "voigt_a": grab("sphere_peel/out.txt", 1, 40, r"voigt_a\s*:\s*" + num),
        "N_HI_pole": grab("sphere_peel/out.txt", 1, 40, r"N\(H  I\)_pole\s*:\s*" + num),
        "tau_pole": grab("sphere_peel/out.txt", 1, 40, r"tau_pole  \(H  I\)\s*:\s*" + num),
        "nphotons": grab("sphere_peel/out.txt", 1, 40, r"Total number of photons\s*:\s*" + num),
        "mean_nscatt": grab("sphere_peel/out.txt", 1, 60, r"Average Number of scattering\s*:\s*" + num),
    },
    "amr_sphere_generic_car_1M": {  # log_car_1M.txt: the Cartesian 64^3 twin of the AMR sphere, T = 1e4 K, tau0 = 1e4
        "voigt_a": grab("amr_sphere_generic/log_car_1M.txt", 1, 40, r"voigt_a\s*:\s*" + num),
        "N_HI_pole": grab("amr_sphere_generic/log_car_1M.txt", 1, 40, r"N\(H  I\)_pole\s*:\s*" + num),
        "tau_pole": grab("amr_sphere_generic/log_car_1M.txt", 1, 40, r"tau_pole  \(H  I\)\s*:\s*" + num),
        "nphotons": grab("amr_sphere_generic/log_car_1M.txt", 1, 40, r"Total number of photons\s*:\s*" + num),
        "mean_nscatt": grab("amr_sphere_generic/log_car_1M.txt", 1, 60, r"Average Number of scattering\s*:\s*" + num),
    },
    "clump_NHI18_fcov1": {  # examples/clump_sphere/log_back, first run
        "N_clumps": grab("clump_sphere/log_back", 1, 56, r"N_clumps\s*=\s*(\d+)", int),
        "f_vol": grab("clump_sphere/log_back", 1, 56, r"f_vol\s*=\s*" + num),
        "cl_rhokap": grab("clump_sphere/log_back", 1, 56, r"cl_rhokap\s*=\s*" + num),
        "cl_Dfreq": grab("clump_sphere/log_back", 1, 56, r"cl_Dfreq\s*=\s*" + num),
        "csr_registrations": grab("clump_sphere/log_back", 1, 56, r"CSR grid:\s*(\d+)\s*registrations", int),
        "csr_cells_per_axis": grab("clump_sphere/log_back", 1, 56, r"registrations in\s*(\d+)\^3", int),
        "tauhomo": grab("clump_sphere/log_back", 1, 56, r"tauhomo\s*=\s*" + num),
        "N_gashomo": grab("clump_sphere/log_back", 1, 56, r"N_gashomo\s*=\s*" + num),
        "cl_vtherm_kms": grab("clump_sphere/log_back", 1, 56, r"cl_vtherm\s*=\s*" + num),
        "nphotons": grab("clump_sphere/log_back", 1, 56, r"Total number of photons\s*:\s*" + num),
        "mean_nscatt": grab("clump_sphere/log_back", 1, 56, r"Average Number of scattering\s*:\s*" + num),
        "wall_minutes": grab("clump_sphere/log_back", 1, 56, r"Total Excution Time\s*:\s*" + num),
    },
}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_logs.json")
json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
print("wrote", dst)
for k, v in out.items():
    print(k, {n: e["value"] for n, e in v.items()})
