// TEST INFRASTRUCTURE: stand-in for <bayesship/dataUtilities.h> (BayesShip is an external library, absent here).
// Only what src/standardPriorLibrary.cpp and src/mcmc_gw_extended.cpp name: the position record handed to probability functions.
#ifndef GWAT_ORACLE_BAYESSHIP_DATAUTILITIES_H
#define GWAT_ORACLE_BAYESSHIP_DATAUTILITIES_H
#include <limits>
namespace bayesship {
static const double limitInf = -std::numeric_limits<double>::infinity();
class positionInfo {
public:
	int dimension = 0;
	bool RJ = false;
	double *parameters = nullptr;
	int *status = nullptr;
	int modelID = 0;
	positionInfo() {}
	positionInfo(int dim, bool rj) : dimension(dim), RJ(rj), owns_(true)
	{
		parameters = new double[dim]();
		if (rj) status = new int[dim]();
	}
	~positionInfo()
	{
		if (owns_) {
			delete[] parameters;
			delete[] status;
		}
	}

private:
	bool owns_ = false;
};
}  // namespace bayesship
#endif
