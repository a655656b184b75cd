// Host-side resolution of GWAT's generation_method strings and detector names.
//
// The reference re-parses the method string with std::string::find for every waveform (src/waveform_generator.cpp:129-275,
// src/ppE_utilities.cpp:65-134, 158-359).  Here it is parsed once per call on the host into a family id (which kernel
// instantiation runs) and a theory id (how the setup kernel derives the ppE betas).
#ifndef GWAT_METHOD_H
#define GWAT_METHOD_H

#include <cstring>
#include <string>

#include "gwat_theory.h"

namespace gwat {

enum FamilyId {
	FAM_D = 0,          // IMRPhenomD
	FAM_D_PPE_INS,      // ppE_IMRPhenomD_Inspiral (+ dCS_/EdGB_IMRPhenomD, which map onto it)
	FAM_D_PPE_IMR,      // ppE_IMRPhenomD_IMR
	FAM_D_GIMR,         // gIMRPhenomD
	FAM_D_NRT,          // IMRPhenomD_NRT
	FAM_D_NRT_PPE_INS,  // ppE_IMRPhenomD_NRT_Inspiral
	FAM_D_NRT_PPE_IMR,  // ppE_IMRPhenomD_NRT_IMR
	FAM_P,              // IMRPhenomPv2 (and IMRPhenomPv2_NRT, which adds no tidal terms: src/IMRPhenomP_NRT.cpp)
	FAM_P_PPE_INS,      // ppE_IMRPhenomPv2_Inspiral
	FAM_P_PPE_IMR,      // ppE_IMRPhenomPv2_IMR
	FAM_P_GIMR,         // gIMRPhenomPv2
	FAM_COUNT
};


struct MethodDesc {
	int family_id;
	int theory;
	bool pv2;
	bool nrt;
	bool ppe;   // betas/b's travel in the parameter vector / source record
	bool gimr;
	bool mcmc;  // the string carried the "MCMC_" prefix (Fisher parameterisation switch, src/fisher.cpp:1845)
	std::string base;  // method with the MCMC_ prefix stripped
};

// 0 on success, -1 if the string names nothing this library implements.
inline int parse_method(const char *method_c, MethodDesc &d)
{
	if (!method_c) return -1;
	std::string m(method_c);
	d.mcmc = false;
	if (m.compare(0, 5, "MCMC_") == 0) {  // local_generation_method, src/fisher.cpp:1754-1768
		d.mcmc = true;
		m.erase(0, 5);
	}
	d.base = m;
	d.theory = THEORY_NONE;
	d.pv2 = m.find("Pv2") != std::string::npos;
	d.nrt = m.find("NRT") != std::string::npos;
	d.ppe = m.find("ppE") != std::string::npos;
	d.gimr = m.find("gIMR") != std::string::npos;

	struct Entry {
		const char *name;
		int fam;
		int theory;
	};
	static const Entry table[] = {
	    {"IMRPhenomD", FAM_D, THEORY_NONE},
	    {"ppE_IMRPhenomD_Inspiral", FAM_D_PPE_INS, THEORY_NONE},
	    {"ppE_IMRPhenomD_IMR", FAM_D_PPE_IMR, THEORY_NONE},
	    {"gIMRPhenomD", FAM_D_GIMR, THEORY_NONE},
	    {"IMRPhenomD_NRT", FAM_D_NRT, THEORY_NONE},
	    {"ppE_IMRPhenomD_NRT_Inspiral", FAM_D_NRT_PPE_INS, THEORY_NONE},
	    {"ppE_IMRPhenomD_NRT_IMR", FAM_D_NRT_PPE_IMR, THEORY_NONE},
	    {"IMRPhenomPv2", FAM_P, THEORY_NONE},
	    {"IMRPhenomPv2_NRT", FAM_P, THEORY_NONE},
	    {"ppE_IMRPhenomPv2_Inspiral", FAM_P_PPE_INS, THEORY_NONE},
	    {"ppE_IMRPhenomPv2_IMR", FAM_P_PPE_IMR, THEORY_NONE},
	    {"gIMRPhenomPv2", FAM_P_GIMR, THEORY_NONE},
	    // theory-mapped methods (assign_mapping, src/ppE_utilities.cpp:158-359): inspiral-only ppE with derived betas
	    {"dCS_IMRPhenomD", FAM_D_PPE_INS, THEORY_DCS},
	    {"EdGB_IMRPhenomD", FAM_D_PPE_INS, THEORY_EDGB},
	    {"dCS_IMRPhenomD_NRT", FAM_D_NRT_PPE_INS, THEORY_DCS},
	    {"EdGB_IMRPhenomD_NRT", FAM_D_NRT_PPE_INS, THEORY_EDGB},
	    {"dCS_IMRPhenomPv2", FAM_P_PPE_INS, THEORY_DCS},
	    {"EdGB_IMRPhenomPv2", FAM_P_PPE_INS, THEORY_EDGB},
	};
	for (const Entry &e : table) {
		if (m == e.name) {
			d.family_id = e.fam;
			d.theory = e.theory;
			return 0;
		}
	}
	// the remaining theory-mapped methods (assign_mapping, src/ppE_utilities.cpp:158-359): "<theory>_<base model>", with
	// "_Inspiral"/"_IMR" after the base model for the two generic re-parameterisations
	struct Theory {
		const char *prefix;
		int id;
		int generic;  // 1: inspiral/IMR chosen by the suffix; 0: inspiral-only ppE; 2: always the IMR ppE
	};
	static const Theory theories[] = {
	    {"EdGB_HO_LO_", THEORY_EDGB_HO_LO, 0}, {"EdGB_HO_", THEORY_EDGB, 0},          {"EdGB_GHOv1_", THEORY_EDGB_GHOV1, 0},
	    {"EdGB_GHOv2_", THEORY_EDGB_GHOV2, 0}, {"EdGB_GHOv3_", THEORY_EDGB_GHOV3, 0}, {"ExtraDimension_", THEORY_EXTRADIM, 0},
	    {"BHEvaporation_", THEORY_BHEVAP, 0},  {"TVG_", THEORY_TVG, 0},               {"DipRad_", THEORY_DIPRAD, 0},
	    {"NonComm_", THEORY_NONCOMM, 0},       {"PNSeries_ppE_", THEORY_PNSERIES, 1}, {"ppEAlt_", THEORY_PPEALT, 1},
	    {"ModDispersion_", THEORY_MODDISP, 2},
	};
	for (const Theory &t : theories) {
		const std::string pre(t.prefix);
		if (m.compare(0, pre.size(), pre) != 0) continue;
		std::string rest = m.substr(pre.size());
		bool ins = t.generic != 2;
		if (t.generic == 1) {
			const std::string a = "_Inspiral", b = "_IMR";
			if (rest.size() > a.size() && rest.compare(rest.size() - a.size(), a.size(), a) == 0) rest.erase(rest.size() - a.size());
			else if (rest.size() > b.size() && rest.compare(rest.size() - b.size(), b.size(), b) == 0) {
				rest.erase(rest.size() - b.size());
				ins = false;
			} else return -1;
		}
		if (rest == "IMRPhenomD") d.family_id = ins ? FAM_D_PPE_INS : FAM_D_PPE_IMR;
		else if (rest == "IMRPhenomD_NRT") d.family_id = ins ? FAM_D_NRT_PPE_INS : FAM_D_NRT_PPE_IMR;
		else if (rest == "IMRPhenomPv2") d.family_id = ins ? FAM_P_PPE_INS : FAM_P_PPE_IMR;
		else return -1;
		d.theory = t.id;
		d.ppe = true;
		return 0;
	}
	return -1;
}

// Row of the generated detector table for a GWAT detector name (src/detector_util.cpp:1083-1158), -1 if unknown.
inline int detector_index(const char *name_c)
{
	if (!name_c) return -1;
	const std::string n(name_c);
	if (n == "Hanford" || n == "hanford") return 0;
	if (n == "Livingston" || n == "livingston") return 1;
	if (n == "Virgo" || n == "virgo") return 2;
	if (n == "Kagra" || n == "kagra") return 3;
	if (n == "Indigo" || n == "indigo") return 4;
	if (n == "Cosmic Explorer" || n == "cosmic explorer" || n == "CE") return 5;
	if (n == "Einstein Telescope 1" || n == "einstein telescope 1" || n == "ET1") return 6;
	if (n == "Einstein Telescope 2" || n == "einstein telescope 2" || n == "ET2") return 7;
	if (n == "Einstein Telescope 3" || n == "einstein telescope 3" || n == "ET3") return 8;
	return -1;
}

}  // namespace gwat
#endif
