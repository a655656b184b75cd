// TEST INFRASTRUCTURE: stand-in for <bayesship/utilities.h>: the CSV writers the reference's example programs call
// (examples/*/src/*.cpp).  Values are written with 17 significant digits (lossless for doubles), one row per line.
#ifndef GWAT_ORACLE_BAYESSHIP_UTILITIES_H
#define GWAT_ORACLE_BAYESSHIP_UTILITIES_H
#include <cstdio>
#include <string>
namespace bayesship {
inline void writeCSVFile(std::string filename, double **data, int rows, int cols)
{
	FILE *fp = std::fopen(filename.c_str(), "w");
	if (!fp) {
		std::fprintf(stderr, "writeCSVFile: cannot open %s\n", filename.c_str());
		return;
	}
	for (int i = 0; i < rows; i++) {
		for (int j = 0; j < cols; j++) std::fprintf(fp, j + 1 < cols ? "%.17g," : "%.17g\n", data[i][j]);
	}
	std::fclose(fp);
}
inline void writeCSVFile(std::string filename, double *data, int n)
{
	FILE *fp = std::fopen(filename.c_str(), "w");
	if (!fp) return;
	for (int j = 0; j < n; j++) std::fprintf(fp, j + 1 < n ? "%.17g," : "%.17g\n", data[j]);
	std::fclose(fp);
}
}  // namespace bayesship
#endif
