#!/usr/bin/env python3
"""Generate gw_analysis_tools_b200/csrc/gwat_tables.inc from the numerical DATA tables of the reference.

The tables are physical constants / published fit coefficients the path cannot exist without:
  * Kerr QNM ringdown/damping frequencies vs final spin (include/gwat/QNM_data.h; originally LALSuite),
  * the 19x11 IMRPhenomD phenomenological fit coefficients of Khan et al. (include/gwat/IMRPhenomD.h:190-270),
  * detector response tensors, vertex locations and geometric factors (include/gwat/detector_util.h:29-157).
They are parsed out of /root/reference at generation time and re-emitted in this project's own layout.  For the QNM
table the natural-cubic-spline second-derivative coefficients are PRECOMPUTED here (the reference re-solves the
1003-point tridiagonal system twice for every waveform, src/IMRPhenomD.cpp:1190-1216) with exactly the operation order
of GSL's cspline_init / solve_tridiag, so that a device-side lookup reproduces gsl_spline_eval bit for bit.

Run:  python tools/gen_tables.py [/root/reference]        (the output is committed; the GPU box never needs the reference)
"""
import os
import re
import sys

_args = [a for a in sys.argv[1:] if not a.startswith("--")]
REF = _args[0] if _args else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gw_analysis_tools_b200", "csrc", "gwat_tables.inc")


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def numbers(body):
    # the reference writes some negative entries as "- 0.12" (sign, blank, digits)
    toks = re.findall(r"[-+]?\s*(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?", body)
    return [float(re.sub(r"\s+", "", t)) for t in toks]


def array_body(text, name):
    m = re.search(re.escape(name) + r"\s*(?:\[[^\]]*\])+\s*=\s*\{(.*?)\}\s*;", text, flags=re.S)
    if not m:
        raise SystemExit("table %s not found" % name)
    return m.group(1).replace("\\\n", " ")


def natural_cspline_c(xa, ya):
    """c[] of GSL's cspline (interpolation/cspline.c + linalg/tridiag.c solve_tridiag), same operation order."""
    n = len(xa)
    c = [0.0] * n
    N = n - 2
    g, diag, off = [0.0] * N, [0.0] * N, [0.0] * N
    for i in range(N):
        h_i = xa[i + 1] - xa[i]
        h_ip1 = xa[i + 2] - xa[i + 1]
        yd_i = ya[i + 1] - ya[i]
        yd_ip1 = ya[i + 2] - ya[i + 1]
        g_i = 1.0 / h_i if h_i != 0.0 else 0.0
        g_ip1 = 1.0 / h_ip1 if h_ip1 != 0.0 else 0.0
        off[i] = h_ip1
        diag[i] = 2.0 * (h_ip1 + h_i)
        g[i] = 3.0 * (yd_ip1 * g_ip1 - yd_i * g_i)
    gamma, alpha, cc, z = [0.0] * N, [0.0] * N, [0.0] * N, [0.0] * N
    alpha[0] = diag[0]
    gamma[0] = off[0] / alpha[0]
    for i in range(1, N - 1):
        alpha[i] = diag[i] - off[i - 1] * gamma[i - 1]
        gamma[i] = off[i] / alpha[i]
    alpha[N - 1] = diag[N - 1] - off[N - 2] * gamma[N - 2]
    z[0] = g[0]
    for i in range(1, N):
        z[i] = g[i] - gamma[i - 1] * z[i - 1]
    for i in range(N):
        cc[i] = z[i] / alpha[i]
    x = [0.0] * N
    x[N - 1] = cc[N - 1]
    for i in range(N - 2, -1, -1):
        x[i] = cc[i] - gamma[i] * x[i + 1]
    for i in range(N):
        c[i + 1] = x[i]
    return c


def main():
    qnm = strip_comments(open(os.path.join(REF, "include/gwat/QNM_data.h")).read())
    a = numbers(array_body(qnm, "QNMData_a"))
    fr = numbers(array_body(qnm, "QNMData_fring"))
    fd = numbers(array_body(qnm, "QNMData_fdamp"))
    n = int(re.search(r"QNMData_length\s*=\s*(\d+)", qnm).group(1))
    assert len(a) == len(fr) == len(fd) == n, (len(a), len(fr), len(fd), n)
    c_fr = natural_cspline_c(a, fr)
    c_fd = natural_cspline_c(a, fd)

    phd = strip_comments(open(os.path.join(REF, "include/gwat/IMRPhenomD.h")).read())
    lam = numbers(array_body(phd, "lambda_num_params"))
    assert len(lam) == 19 * 11, len(lam)

    det = strip_comments(open(os.path.join(REF, "include/gwat/detector_util.h")).read())
    dets = [("Hanford", "H", "Hanford_D"), ("Livingston", "L", "Livingston_D"), ("Virgo", "V", "Virgo_D"),
            ("Kagra", "K", "Kagra_D"), ("Indigo", "I", "Indigo_D"), ("CE", "CE", "CE_D"),
            ("ET1", "ET1", "ET1_D"), ("ET2", "ET2", "ET2_D"), ("ET3", "ET3", "ET3_D")]
    rows = []
    for name, pre, dname in dets:
        D = numbers(array_body(det, dname))
        loc = numbers(array_body(det, pre + "_location"))
        gf = float(re.search(pre + r"_geometric_factor\s*=\s*([-+0-9.eE]+)", det).group(1))
        assert len(D) == 9 and len(loc) == 3
        rows.append((name, D, loc, gf))

    # redshift(luminosity distance) piecewise half-power series, all cosmologies of the reference
    # (include/gwat/D_Z_Config.h: cosmos, boundaries_D, COEFF_VEC_DZ; evaluated by Z_from_DL, src/util.cpp:356-382)
    dz = strip_comments(open(os.path.join(REF, "include/gwat/D_Z_Config.h")).read())
    bD = numbers(array_body(dz, "boundaries_D"))
    cDZ = numbers(array_body(dz, "COEFF_VEC_DZ"))
    ncos = int(re.search(r"num_cosmologies\s*=\s*(\d+)", dz).group(1))
    nseg, ndeg = 3, 12
    assert len(bD) == ncos * (nseg + 1) and len(cDZ) == ncos * nseg * ndeg, (len(bD), len(cDZ))
    cosmo_names = re.findall(r'"([A-Z0-9_]+)"', re.search(r"cosmos\[\d+\]\s*=\s*\{([^}]*)\}", dz).group(1))
    assert len(cosmo_names) == ncos and cosmo_names[0] == "PLANCK15", cosmo_names

    # modified-dispersion distance D_alpha(z) (include/gwat/D_Z_Config_modified_dispersion.h: MD_alphas, MD_boundaries_Z,
    # MD_COEFF_VEC_ZD; evaluated by DL_from_Z_MD, src/ppE_utilities.cpp:785-822: sum_j c_j z^(-3.5 + j/2))
    md = strip_comments(open(os.path.join(REF, "include/gwat/D_Z_Config_modified_dispersion.h")).read())
    md_alphas = numbers(array_body(md, "MD_alphas"))
    md_bz = numbers(array_body(md, "MD_boundaries_Z"))
    md_c = numbers(array_body(md, "MD_COEFF_VEC_ZD"))
    na, md_seg, md_deg = len(md_alphas), 3, 17
    assert na == 9 and len(md_bz) == na * (md_seg + 1) and len(md_c) == na * md_seg * md_deg, (na, len(md_bz), len(md_c))

    r = repr
    with open(OUT, "w") as o:
        o.write("// GENERATED by tools/gen_tables.py from the reference's numerical data tables -- do not edit.\n")
        o.write("// Layout is this project's own; see the generator for provenance (file:line) of every table.\n\n")
        o.write("#define GWAT_QNM_N %d\n" % n)
        o.write("// {a_final, M*f_ring, spline c (f_ring), M*f_damp, spline c (f_damp)} -- natural cubic spline, c = y''/2\n")
        o.write("GWAT_TABLE_QUALIFIER double gwat_qnm_knots[GWAT_QNM_N][5] = {\n")
        for i in range(n):
            o.write("{%s,%s,%s,%s,%s},\n" % (r(a[i]), r(fr[i]), r(c_fr[i]), r(fd[i]), r(c_fd[i])))
        o.write("};\n\n")
        o.write("// IMRPhenomD fit coefficients, rows: rho1-3, v2, gamma1-3, sigma1-4, beta1-3, alpha1-5 (arXiv:1508.07253 tab. V)\n")
        o.write("GWAT_TABLE_QUALIFIER double gwat_phenomd_fit[19][11] = {\n")
        for i in range(19):
            o.write("{" + ",".join(r(x) for x in lam[i * 11:(i + 1) * 11]) + "},\n")
        o.write("};\n\n")
        o.write("// z(D_L/Mpc) per cosmology (index = position in the reference's cosmos[]: " + ", ".join(cosmo_names) + "):\n")
        o.write("// segment boundaries and, per segment, coefficients of sum_k c_k (sqrt D_L)^k\n")
        o.write("#define GWAT_NUM_COSMOLOGIES %d\n#define GWAT_DZ_SEGMENTS %d\n#define GWAT_DZ_DEGREE %d\n" % (ncos, nseg, ndeg))
        o.write("#define GWAT_COSMOLOGY_NAMES {" + ",".join('"%s"' % n for n in cosmo_names) + "}\n")
        o.write("GWAT_TABLE_QUALIFIER double gwat_dz_boundaries[GWAT_NUM_COSMOLOGIES][GWAT_DZ_SEGMENTS + 1] = {\n")
        for c in range(ncos):
            o.write("{" + ",".join(r(x) for x in bD[c * (nseg + 1):(c + 1) * (nseg + 1)]) + "},\n")
        o.write("};\n")
        o.write("GWAT_TABLE_QUALIFIER double gwat_dz_coeffs[GWAT_NUM_COSMOLOGIES][GWAT_DZ_SEGMENTS][GWAT_DZ_DEGREE] = {\n")
        for c in range(ncos):
            o.write("{")
            for i in range(nseg):
                base = (c * nseg + i) * ndeg
                o.write("{" + ",".join(r(x) for x in cDZ[base:base + ndeg]) + "},")
            o.write("},\n")
        o.write("};\n\n")
        o.write("// modified-dispersion distance D_alpha(z)/Mpc: alphas, z boundaries per alpha, coefficients of sum_j c_j z^(-3.5 + j/2)\n")
        o.write("#define GWAT_MD_ALPHAS %d\n#define GWAT_MD_SEGMENTS %d\n#define GWAT_MD_DEGREE %d\n" % (na, md_seg, md_deg))
        o.write("GWAT_TABLE_QUALIFIER double gwat_md_alphas[GWAT_MD_ALPHAS] = {" + ",".join(r(x) for x in md_alphas) + "};\n")
        o.write("GWAT_TABLE_QUALIFIER double gwat_md_boundaries_z[GWAT_MD_ALPHAS][GWAT_MD_SEGMENTS + 1] = {\n")
        for i in range(na):
            o.write("{" + ",".join(r(x) for x in md_bz[i * (md_seg + 1):(i + 1) * (md_seg + 1)]) + "},\n")
        o.write("};\n")
        o.write("GWAT_TABLE_QUALIFIER double gwat_md_coeffs[GWAT_MD_ALPHAS][GWAT_MD_SEGMENTS][GWAT_MD_DEGREE] = {\n")
        for i in range(na):
            o.write("{")
            for k in range(md_seg):
                base = (i * md_seg + k) * md_deg
                o.write("{" + ",".join(r(x) for x in md_c[base:base + md_deg]) + "},")
            o.write("},\n")
        o.write("};\n\n")
        o.write("#define GWAT_NUM_KNOWN_DETECTORS %d\n" % len(rows))
        o.write("// per detector: 9 response-tensor entries (row-major), 3 vertex coordinates [m], geometric factor\n")
        o.write("GWAT_TABLE_QUALIFIER double gwat_detector_table[GWAT_NUM_KNOWN_DETECTORS][13] = {\n")
        for name, D, loc, gf in rows:
            o.write("/* %s */ {" % name + ",".join(r(x) for x in D + loc + [gf]) + "},\n")
        o.write("};\n")
    print("wrote", os.path.normpath(OUT))


def write_zd_tables():
    """D_L(z) for the gwatpy helper DL_from_Z_py (host code only, so a file of its own: regenerating it does not touch the kernels'
    tables): include/gwat/D_Z_Config.h boundaries_Z and COEFF_VEC_ZD, evaluated by DL_from_Z (src/util.cpp:422-450)."""
    dz = strip_comments(open(os.path.join(REF, "include/gwat/D_Z_Config.h")).read())
    bZ = numbers(array_body(dz, "boundaries_Z"))
    cZD = numbers(array_body(dz, "COEFF_VEC_ZD"))
    ncos = int(re.search(r"num_cosmologies\s*=\s*(\d+)", dz).group(1))
    nseg, ndeg = 3, 12
    assert len(bZ) == ncos * (nseg + 1) and len(cZD) == ncos * nseg * ndeg, (len(bZ), len(cZD))
    out = os.path.join(os.path.dirname(OUT), "gwat_tables_zd.inc")
    r = repr
    with open(out, "w") as o:
        o.write("// GENERATED by tools/gen_tables.py (write_zd_tables) from the reference's include/gwat/D_Z_Config.h -- do not edit.\n")
        o.write("// D_L(z)/Mpc per cosmology: segment boundaries in z and, per segment, coefficients of sum_k c_k (sqrt z)^k\n")
        o.write("static const double gwat_zd_boundaries[%d][%d] = {\n" % (ncos, nseg + 1))
        for c in range(ncos):
            o.write("{" + ",".join(r(x) for x in bZ[c * (nseg + 1):(c + 1) * (nseg + 1)]) + "},\n")
        o.write("};\n")
        o.write("static const double gwat_zd_coeffs[%d][%d][%d] = {\n" % (ncos, nseg, ndeg))
        for c in range(ncos):
            o.write("{")
            for i in range(nseg):
                base = (c * nseg + i) * ndeg
                o.write("{" + ",".join(r(x) for x in cZD[base:base + ndeg]) + "},")
            o.write("},\n")
        o.write("};\n")
    print("wrote", os.path.normpath(out))


def write_site_tables():
    """Latitude / longitude of the detector sites for the gwatpy helper get_detector_parameters (host code only): include/gwat/detector_util.h
    *_LAT / *_LONG, rows in the order of gwat_detector_table."""
    det = strip_comments(open(os.path.join(REF, "include/gwat/detector_util.h")).read())
    out = os.path.join(os.path.dirname(OUT), "gwat_tables_sites.inc")
    with open(out, "w") as o:
        o.write("// GENERATED by tools/gen_tables.py (write_site_tables) from the reference's include/gwat/detector_util.h -- do not edit.\n")
        o.write("// per detector (rows as gwat_detector_table): latitude, longitude [rad]\n")
        o.write("static const double gwat_site_lat_long[9][2] = {\n")
        for name, pre in [("Hanford", "H"), ("Livingston", "L"), ("Virgo", "V"), ("Kagra", "K"), ("Indigo", "I"), ("CE", "CE"),
                          ("ET1", "ET1"), ("ET2", "ET2"), ("ET3", "ET3")]:
            lat = float(re.search(r"\b" + pre + r"_LAT\s*=\s*([-+0-9.eE]+)", det).group(1))
            lon = float(re.search(r"\b" + pre + r"_LONG\s*=\s*([-+0-9.eE]+)", det).group(1))
            o.write("/* %s */ {%r,%r},\n" % (name, lat, lon))
        o.write("};\n")
    print("wrote", os.path.normpath(out))


if __name__ == "__main__":
    if "--zd-only" in sys.argv:
        write_zd_tables()
    elif "--sites-only" in sys.argv:
        write_site_tables()
    else:
        write_site_tables()
        main()
        write_zd_tables()
