// Detector noise curves: the input-preparation side of the likelihood path.
//
// populate_noise (src/detector_util.cpp:87-282) returns the amplitude spectral density sqrt(S_n(f)) of a named curve:
//   * analytic models  aLIGO_analytic (:288-295), Hanford_O1_fitted (:412-419)           -- evaluated directly
//   * tabulated curves (AdLIGODesign, AdLIGOAPlus, CE1/2, AdVIRGOPlus*, KAGRA_*, ET-D, AdLIGOVoyager, ...): two-column CSV
//     files of (f, sqrt S) shipped under data/noise_data/currently_supported, interpolated linearly (gsl_interp_linear)
// The LISA curves are outside this path (space detector response is out of scope).
// This is host code that runs once per analysis, before gwat_b200_set_network; it reads the reference's own CSV files from
// a directory the caller names (GWAT installs them under GWAT_SHARE_DIR/noise_data) -- none are copied into this repository.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gwat_b200.h"

namespace {

// curve name -> file name, as populate_noise maps them (src/detector_util.cpp:166-254)
struct CurveFile {
	const char *curve, *file;
};
const CurveFile kCurveFiles[] = {
    {"AdLIGOMidHigh", "AdLIGOMidHigh.csv"},
    {"_AdLIGODesign", "AdLIGODesign.csv"},
    {"AdLIGODesign", "aligo_O4high.csv"},
    {"AdLIGODesign_smoothed", "aligo_O4high_smoothed.csv"},
    {"AdLIGOAPlus", "AplusDesign.csv"},
    {"AdLIGOAPlus_smoothed", "AplusDesign_smoothed.csv"},
    {"CE1", "CE1_strain.csv"},
    {"CE1_smoothed", "CE1_strain_smoothed.csv"},
    {"CE2", "CE2_strain.csv"},
    {"CE2_smoothed", "CE2_strain_smoothed.csv"},
    {"AdVIRGOPlus2_opt", "avirgo_O5high_NEW.csv"},
    {"AdVIRGOPlus2_opt_smoothed", "avirgo_O5high_NEW_smoothed.csv"},
    {"AdVIRGOPlus2_pess", "avirgo_O5low_NEW.csv"},
    {"AdVIRGOPlus2_pess_smoothed", "avirgo_O5low_NEW_smoothed.csv"},
    {"AdVIRGOPlus1", "avirgo_O4high_NEW.csv"},
    {"AdVIRGOPlus1_smoothed", "avirgo_O4high_NEW_smoothed.csv"},
    {"KAGRA_opt", "kagra_128Mpc.csv"},
    {"KAGRA_pess", "kagra_80Mpc.csv"},
    {"ET-D", "ET-0000A-18_ETDSensitivityCurveTxtFile.csv"},
    {"ET-D_smoothed", "ETDXylophoneDwyer.csv"},
    {"AdLIGOVoyager", "Voyager.csv"},
};

double aligo_analytic(double f)
{
	const double S = 3e-48, fknee = 70.;
	const double x = fknee / f;
	const double x4 = x * x * x * x;
	return std::sqrt(S * (x4 + 2 + 2 * x * x) / 5);
}

double hanford_o1_fitted(double f)
{
	const double a[7] = {47.8466, -92.1896, 35.9273, -7.61447, 0.916742, -0.0588089, 0.00156345};
	const double S0 = .8464;
	const double x = std::log(f);
	return std::sqrt(S0) * std::exp(a[0] + a[1] * x + a[2] * x * x + a[3] * x * x * x + a[4] * x * x * x * x + a[5] * x * x * x * x * x +
	                                a[6] * x * x * x * x * x * x);
}

// Rows of "a , b" (the reference's read_file, src/io_util.cpp: comma-separated doubles, one row per line).
bool read_two_columns(const std::string &path, std::vector<double> &x, std::vector<double> &y)
{
	FILE *fp = std::fopen(path.c_str(), "r");
	if (!fp) return false;
	char line[512];
	while (std::fgets(line, sizeof line, fp)) {
		double a, b;
		if (std::sscanf(line, " %lf , %lf", &a, &b) == 2 || std::sscanf(line, " %lf %lf", &a, &b) == 2) {
			x.push_back(a);
			y.push_back(b);
		}
	}
	std::fclose(fp);
	return x.size() >= 2;
}

}  // namespace

extern "C" int gwat_b200_populate_noise(const double *frequencies, const char *curve, const char *noise_data_dir, int length,
                                        double *noise_root)
{
	if (!frequencies || !curve || !noise_root || length < 0) return GWAT_B200_ERR_ARG;
	const std::string name(curve);
	if (name == "aLIGO_analytic") {
		for (int i = 0; i < length; i++) noise_root[i] = aligo_analytic(frequencies[i]);
		return GWAT_B200_OK;
	}
	if (name == "Hanford_O1_fitted") {
		for (int i = 0; i < length; i++) noise_root[i] = hanford_o1_fitted(frequencies[i]);
		return GWAT_B200_OK;
	}
	if (name.compare(0, 4, "LISA") == 0) return GWAT_B200_ERR_UNSUPPORTED;
	const char *file = nullptr;
	for (const CurveFile &c : kCurveFiles)
		if (name == c.curve) file = c.file;
	if (!file) return GWAT_B200_ERR_ARG;  // the reference prints "Detector ... not supported" and leaves the output untouched
	if (!noise_data_dir) return GWAT_B200_ERR_ARG;
	std::string path(noise_data_dir);
	if (!path.empty() && path.back() != '/') path += '/';
	std::vector<double> x, y;
	if (!read_two_columns(path + file, x, y)) return GWAT_B200_ERR_STATE;
	// gsl_interp_linear: y_lo + (x - x_lo) / (x_hi - x_lo) * (y_hi - y_lo) on the interval found by bisection; outside the
	// table GSL raises a domain error (abort() under its default handler) -- here NaN and an error code.
	const size_t n = x.size();
	int rc = GWAT_B200_OK;
	size_t lo = 0;  // gsl_interp_accel: the last interval is tried first
	for (int i = 0; i < length; i++) {
		const double f = frequencies[i];
		if (!(f >= x[0] && f <= x[n - 1])) {
			noise_root[i] = NAN;
			rc = GWAT_B200_ERR_ARG;
			continue;
		}
		if (!(f >= x[lo] && f < x[lo + 1])) {
			size_t a = 0, b = n - 1;
			while (b > a + 1) {
				const size_t m = (a + b) / 2;
				if (x[m] > f) b = m;
				else a = m;
			}
			lo = a;
		}
		const double dx = x[lo + 1] - x[lo];
		noise_root[i] = dx > 0.0 ? y[lo] + (f - x[lo]) / dx * (y[lo + 1] - y[lo]) : 0.0;  // (GSL gives 0 for a repeated knot)
	}
	return rc;
}

// gps_to_GMST_radian (src/util.cpp:1793-1846): Greenwich mean sidereal time of a GPS time by the USNO approximation the
// reference uses (GPS -> Julian date without leap seconds, GMST = 6.697374558 + 0.06570982441908 D0 + 1.00273790935 H +
// 0.000026 tau^2 hours, reduced modulo 24), in radians.  It is how callers obtain the `gmst` every likelihood call takes.
extern "C" double gwat_b200_gps_to_gmst_radian(double gps_time)
{
	const double J2000 = 2451545;
	const double JD = J2000 + (gps_time - 630763213.) / (86400.);
	const double JD0 = std::floor(JD) + .5;
	const double H = (JD - JD0) * 24;
	const double D0 = JD0 - J2000, Dd = JD - J2000;
	const double tau = Dd / 36525.;
	const double unscaled = 6.697374558 + 0.06570982441908 * D0 + 1.00273790935 * H + 0.000026 * tau * tau;
	const double hours = ((int)std::floor(unscaled) % 24);
	return (hours + (unscaled - std::floor(unscaled))) * M_PI / 12.;
}
