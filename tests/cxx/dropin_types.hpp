// TEST ONLY.  A stand-in for the members of gen_params_base<double> (reference include/gwat/util.h:118-290) that the hot path
// reads, with the reference's member names, types (bool flags, pointer-valued modification arrays) and defaults, so that
// include/gwat_b200_cxx.hpp can be exercised on a box that has no GWAT headers.  When the real headers are present
// (-DGWAT_CXX_REAL_HEADERS, used by tests/test_cxx_adapter.py in the build container) the shim below is compiled against
// gen_params_base<double> itself instead.
#ifndef GWAT_B200_TEST_DROPIN_TYPES_HPP
#define GWAT_B200_TEST_DROPIN_TYPES_HPP
#ifdef GWAT_CXX_REAL_HEADERS
#include <util.h>
typedef gen_params_base<double> test_gen_params;
#else
#include <cstddef>
#include <string>
struct test_gen_params {
	std::string cosmology = "PLANCK15";
	double mass1, mass2, Luminosity_Distance;
	double spin1[3], spin2[3];
	double tc = 0;
	double tidal1 = -1, tidal2 = -1, tidal_s = -1, tidal_a = -1, tidal_weighted = -1;
	bool tidal_love = true, tidal_love_error = false;
	double delta_tidal_weighted = -1;
	double diss_tidal1 = -1, diss_tidal2 = -1, diss_tidal_s = -1, diss_tidal_a = -1, diss_tidal_weighted = -1;
	double psi = 0, incl_angle;
	bool equatorial_orientation = false;
	double theta_l, phi_l;
	bool horizon_coord = false;
	double theta, phi, RA, DEC;
	double gmst;
	bool NSflag1 = false, NSflag2 = false, dep_postmerger = false;
	double f_ref = 0;
	bool shift_time = true, shift_phase = true;
	double phiRef = 0;
	bool sky_average = false;
	double chip = -1, phip = -1;
	int Nmod_beta = 0, Nmod_alpha = 0, Nmod_sigma = 0, Nmod_phi = 0;
	int *betai = NULL, *alphai = NULL, *sigmai = NULL, *phii = NULL;
	double *delta_beta = NULL, *delta_alpha = NULL, *delta_sigma = NULL, *delta_phi = NULL;
	double *bppe = NULL;
	double *betappe = NULL;
	int Nmod = 0;
	int PNorder = 35;
};
#endif
#endif
