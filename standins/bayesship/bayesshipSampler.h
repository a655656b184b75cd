// TEST INFRASTRUCTURE: stand-in for <bayesship/bayesshipSampler.h>: the abstract probability function the reference's priors
// and likelihood wrappers derive from (src/standardPriorLibrary.cpp, src/mcmc_gw_extended.cpp:513-518).
#ifndef GWAT_ORACLE_BAYESSHIP_SAMPLER_H
#define GWAT_ORACLE_BAYESSHIP_SAMPLER_H
#include "bayesship/dataUtilities.h"
namespace bayesship {
class probabilityFn {
public:
	virtual ~probabilityFn() {}
	virtual double eval(positionInfo *position, int chainID) { return 0; }
};
}  // namespace bayesship
#endif
