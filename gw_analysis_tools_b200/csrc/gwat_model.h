// Per-walker data structures shared by the setup and the per-bin code of the gwat_b200 kernels.
#ifndef GWAT_MODEL_H
#define GWAT_MODEL_H

#include "gwat_hd.h"
#include "../../include/gwat_b200.h"

namespace gwat {

// Physical constants with the reference's values (include/gwat/util.h:44-58).
#define GWAT_MSOL_SEC 4.925491025543575903411922162094833998e-6
#define GWAT_C_SI 299792458.
#define GWAT_MPC_SEC (3.085677581491367278913937957796471611e22 / GWAT_C_SI)
#define GWAT_GAMMA_E 0.5772156649015328606065120900824024310421

// ---- waveform family, resolved on the host from the generation_method string ------------------------------------------
// The reference dispatches on std::string::find at run time (src/waveform_generator.cpp:129-275) and through virtual
// overrides; here the combination is a compile-time tag so every family gets its own specialised kernel.
enum Base { BASE_D = 0, BASE_P = 1 };          // IMRPhenomD carrier, or IMRPhenomPv2 twist-up of it
enum PpeMode { PPE_NONE = 0, PPE_INSPIRAL = 1, PPE_IMR = 2 };

template <int BASE_, int PPE_, bool GIMR_, bool NRT_>
struct Family {
	static constexpr int base = BASE_;
	static constexpr int ppe = PPE_;
	static constexpr bool gimr = GIMR_;
	static constexpr bool nrt = NRT_;
};

// How the ppE betas of a walker are produced in the setup kernel (theory mapping, src/ppE_utilities.cpp:158-359).
enum Theory {
	TH_NONE = 0,  // betas are given directly (ppE_*) or no ppE terms
	TH_DCS = 1,
	TH_EDGB = 2,
};

// Derived per-walker quantities in seconds -- what the reference keeps in source_parameters<double>
// (include/gwat/util.h:426-665), filled by populate_source_parameters (src/util.cpp:997-1028).
struct SrcQ {
	double mass1, mass2, M, q, chirpmass, eta, delta_mass;
	double spin1x, spin1y, spin1z, spin2x, spin2y, spin2z;
	double chi_s, chi_a, chi_eff, chi_pn;
	double DL, A0;
	double tc, phiRef, f_ref, incl_angle;
	double fRD, fdamp, f1, f3, f1_phase, f2_phase;
	bool shift_time, shift_phase, sky_average, dep_postmerger, NSflag1, NSflag2;
	// PhenomPv2
	double chil, chip, phip, SP, SL, s, thetaJN, alpha0, phi_aligned, zeta_polariz;
	// NRT
	double tidal1, tidal2, tidal_weighted, delta_tidal_weighted, diss_tidal_weighted;
	double quad1, quad2, oct1, oct2;
	// ppE / gIMR
	int cosmology;  // index into the reference's cosmos[] (theory mappings: Z_from_DL)
	int Nmod;
	double betappe[GWAT_B200_MAX_MOD], bppe[GWAT_B200_MAX_MOD];
	int Nmod_phi, Nmod_sigma, Nmod_beta, Nmod_alpha;
	int phii[GWAT_B200_MAX_MOD], sigmai[GWAT_B200_MAX_MOD], betai[GWAT_B200_MAX_MOD], alphai[GWAT_B200_MAX_MOD];
	double delta_phi[GWAT_B200_MAX_MOD], delta_sigma[GWAT_B200_MAX_MOD], delta_beta[GWAT_B200_MAX_MOD],
	    delta_alpha[GWAT_B200_MAX_MOD];
};

// The 19 phenomenological coefficients (reference: lambda_parameters<T>, include/gwat/IMRPhenomD.h:17-27); index 0 of
// sigma/beta/alpha holds the connection coefficients, as in the reference.
struct Lambda {
	double rho[3], v2, gamma[3], sigma[5], beta[4], alpha[6];
};

// Everything the per-bin code needs about one walker's IMRPhenomD carrier.  Filled once per walker by the setup kernel,
// read (broadcast) by every bin of that walker.
struct DCoef {
	// region logic
	double fcut;      // 0.2/M: above it the waveform is exactly 0
	double f1a, f3a;  // amplitude: inspiral | intermediate | merger-ringdown
	double f1p, f2p;  // phase:     inspiral | intermediate | merger-ringdown
	// frequency scaling
	double M;
	double sM_hi, sM_lo;  // M^(fl(1/6)) as a double-double
	double logM, logpiM;  // ln M, ln(pi M)
	double pichirp;       // pi * chirpmass (ppE terms)
	// amplitude
	double A0;       // A0 * M^(7/6)
	double ains[7];  // TaylorF2 amplitude coefficients times pi^(k/3)
	double rho[3];
	double ic[5], ix1, ix2, ix3;  // intermediate amplitude: Newton form on nodes [x1,x1,x2,x3,x3] in x = M f
	double mr_num, mr_rate, mr_w2, fRD, fdamp, inv_fdamp;
	// inspiral phase
	double k1, k2, k3, k4, k7;  // static TaylorF2 phase coefficients times pi^(k/3) (k3 also times M)
	double c8, c9, c10, c11;    // log-dependent 2.5PN / 3PN pieces
	double pi53, pi2;           // pi^(5/3), pi^2 (applied per bin, after the log)
	double tf2;                 // 3/(128 eta) * pi^(-5/3)
	double k128;                // 3/(128 eta)
	double inv_eta;
	double sig1M, sig2q, sig3q, sig4q;
	// intermediate and merger-ringdown phase
	double beta0, beta1, beta2, beta3_3;
	double alpha0, alpha1, alpha2, alpha3_43, alpha4, alpha5fRD;
	// time and phase reference
	double tc, f_ref, phic;
	double tc_shift;  // tc = 2 pi t_c + tc_shift (kept so that a stencil point can be re-timed per detector, see k_fisher_setup)
	// ppE / gIMR extras (used only by the families that have them)
	int Nmod;
	double betappe[GWAT_B200_MAX_MOD], bppe[GWAT_B200_MAX_MOD];
	int bint[GWAT_B200_MAX_MOD];  // bppe as an integer when it is one (|b| <= 32), else INT_MIN-like sentinel 9999
	double ppe_scale;             // (pi Mc / M)^(1/3): (pi Mc f)^(1/3) = ppe_scale * (M f)^(1/3)
	int n_gimr_neg;
	double gimr_neg_coef[4];
	int gimr_neg_pow[4];
	// NRT extras
	double nrt_phase_coeff, nrt_ss_coeff, nrt_amp_coeff, nrt_diss_coeff;
	double nrt_fmerger, nrt_fmerger12;
};

struct cplx {
	double re, im;
};

// Per-walker, per-detector projection constants.
struct DetCoef {
	double Fplus, Fcross;  // antenna patterns (detector_response_functions_equatorial, src/detector_util.cpp:900)
	double tshift;         // -2 pi * DTOA: the response is multiplied by exp(i * tshift * f) (src/waveform_util.cpp:173-178)
	// fused-likelihood constants (gwat_like.h): the response is  carrier * (ga * P + gb * Q)  with walker constants
	//   IMRPhenomD families: P = 1, Q = -i      ga = F+ (1+cos^2 i)/2,  gb = Fx cos i
	//   IMRPhenomPv2:        P = h+ twist factor, Q = hx twist factor (before the 2 zeta rotation)
	//                        ga = F+ cos 2z - Fx sin 2z,  gb = F+ sin 2z + Fx cos 2z
	double ga, gb;
};

// PhenomPv2 twist-up constants of one walker.
struct PCoef {
	double A0;  // rescaled amplitude (src/IMRPhenomP.cpp:258)
	double SP, SL, eta;
	double lc1, lc2;  // L2PN series coefficients 1.5 + eta/6, 3.375 - 19 eta/8 - eta^2/24
	double acoef[5], ecoef[5];
	double alpha_const;    // alpha0 - alpha_offset
	double epsilon_offset;
	double Y[5];  // -2Y_{2m}(thetaJN, 0), m = -2..2 (real at zero azimuth)
	// the same harmonics in the combinations the fused likelihood uses (gwat_like.h: carrier_terms):
	//   tw = { Y1 - Y-1, Y1 + Y-1, (sqrt6 / 2) Y0, Y2 + Y-2, Y2 - Y-2, (Y2 + Y-2) / 2, (Y2 - Y-2) / 2 },  SP2 = SP^2
	double tw[7], SP2;
	double c2z, s2z;        // polarisation rotation by 2 zeta (src/waveform_generator.cpp:257-266)
	double phic, tc, f_ref, tcorr_2pi;
};

struct WalkerCoef {
	DCoef d;
	PCoef p;
	double pfac, cfac;  // PhenomD family: (1+cos^2 iota)/2 and cos iota (src/waveform_generator.cpp:196-199)
	DetCoef det[GWAT_B200_MAX_DETECTORS];
	int valid;  // 0 when the parameter point is unphysical (NaN somewhere in the setup): logL = NaN
	int pad_;
};

}  // namespace gwat
#endif
