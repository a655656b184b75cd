// Chain output of the ensemble sampler (host code): the data dump and the thinned, flattened sample file of the reference's
// mcmc_sampler_output (src/mcmc_io_util.cpp:555-990), with the SAME dataset names, shapes and thinning rule.
//
// The reference writes HDF5 (H5Cpp, gzip-6 chunks).  HDF5 is not available where this library is built, so the datasets go into a
// flat, self-describing container instead (a converter to HDF5 is a dozen lines of h5py: every record carries its full path):
//
//   file   := magic "GWATDUMP" u32 version (1) u32 n_records  record*
//   record := u32 path_len  path bytes (no terminator, '/'-separated like an HDF5 path)  u32 dtype (0 = float64, 1 = int32)
//             u32 rank  u64 dims[rank]  payload (row-major, little-endian, dims product * itemsize bytes)
//
// Datasets (src/mcmc_io_util.cpp:689-900): "/MCMC_OUTPUT/CHAIN <id>" [steps][dimension], "/MCMC_OUTPUT/LOGL_LOGP/CHAIN <id>"
// [steps][2], "/MCMC_METADATA/CHAIN TEMPERATURES" [chains], "/MCMC_METADATA/SUGGESTED TRIM LENGTHS" [chains],
// "/MCMC_METADATA/AC VALUES" [cold chains][dimension]; thinned file: "/THINNED_MCMC_OUTPUT/THINNED FLATTENED CHAINS".
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gwat_b200.h"
#include "../../include/gwat_b200_sampler.h"

struct gwat_b200_dump {
	std::FILE *f = nullptr;
	uint32_t n_records = 0;
};

namespace {
const char kMagic[8] = {'G', 'W', 'A', 'T', 'D', 'U', 'M', 'P'};
bool put(std::FILE *f, const void *p, size_t n) { return std::fwrite(p, 1, n, f) == n; }
}  // namespace

extern "C" {

int gwat_b200_dump_create(const char *path, gwat_b200_dump **out)
{
	if (!path || !out) return GWAT_B200_ERR_ARG;
	*out = nullptr;
	std::FILE *f = std::fopen(path, "wb");
	if (!f) return GWAT_B200_ERR_ARG;
	const uint32_t version = 1, zero = 0;
	if (!put(f, kMagic, 8) || !put(f, &version, 4) || !put(f, &zero, 4)) {
		std::fclose(f);
		return GWAT_B200_ERR_ARG;
	}
	gwat_b200_dump *d = new gwat_b200_dump;
	d->f = f;
	*out = d;
	return GWAT_B200_OK;
}

int gwat_b200_dump_write(gwat_b200_dump *d, const char *dataset_path, int dtype, int rank, const long long *dims, const void *data)
{
	if (!d || !d->f || !dataset_path || (dtype != 0 && dtype != 1) || rank < 0 || rank > 8 || (rank > 0 && !dims)) return GWAT_B200_ERR_ARG;
	uint64_t count = 1;
	for (int i = 0; i < rank; i++) {
		if (dims[i] < 0) return GWAT_B200_ERR_ARG;
		count *= (uint64_t)dims[i];
	}
	if (count > 0 && !data) return GWAT_B200_ERR_ARG;
	const uint32_t len = (uint32_t)std::strlen(dataset_path), dt = (uint32_t)dtype, rk = (uint32_t)rank;
	bool ok = put(d->f, &len, 4) && put(d->f, dataset_path, len) && put(d->f, &dt, 4) && put(d->f, &rk, 4);
	for (int i = 0; ok && i < rank; i++) {
		const uint64_t v = (uint64_t)dims[i];
		ok = put(d->f, &v, 8);
	}
	ok = ok && (count == 0 || put(d->f, data, count * (dtype == 0 ? 8 : 4)));
	if (!ok) return GWAT_B200_ERR_ARG;
	d->n_records++;
	return GWAT_B200_OK;
}

int gwat_b200_dump_close(gwat_b200_dump *d)
{
	if (!d) return GWAT_B200_ERR_ARG;
	bool ok = true;
	if (d->f) {
		ok = std::fseek(d->f, 12, SEEK_SET) == 0 && put(d->f, &d->n_records, 4);
		ok = std::fclose(d->f) == 0 && ok;
	}
	delete d;
	return ok ? GWAT_B200_OK : GWAT_B200_ERR_ARG;
}

// create_data_dump (src/mcmc_io_util.cpp:643-990) for chains of one common length (the fixed-ladder sampler of this library):
// positions[n_chains][steps][dimension]; logl_logp[n_chains][steps][2] or NULL; chain_ids / temperatures [n_chains];
// trim_lengths [n_chains] or NULL (zeros); ac_values [n_cold][dimension] (int) or NULL (dataset omitted, as before calc_ac_vals).
int gwat_b200_write_data_dump(const char *path, int n_chains, int dimension, long long steps, const int *chain_ids, const double *temperatures,
                              const double *positions, const double *logl_logp, const int *trim_lengths, int n_cold, const int *ac_values)
{
	if (!path || n_chains <= 0 || dimension <= 0 || steps < 0 || !chain_ids || !temperatures || !positions) return GWAT_B200_ERR_ARG;
	gwat_b200_dump *d = nullptr;
	if (int rc = gwat_b200_dump_create(path, &d)) return rc;
	int rc = GWAT_B200_OK;
	for (int c = 0; c < n_chains && rc == 0; c++) {
		const std::string name = "/MCMC_OUTPUT/CHAIN " + std::to_string(chain_ids[c]);
		const long long dims[2] = {steps, dimension};
		rc = gwat_b200_dump_write(d, name.c_str(), 0, 2, dims, positions + (size_t)c * steps * dimension);
		if (rc == 0 && logl_logp) {
			const std::string n2 = "/MCMC_OUTPUT/LOGL_LOGP/CHAIN " + std::to_string(chain_ids[c]);
			const long long d2[2] = {steps, 2};
			rc = gwat_b200_dump_write(d, n2.c_str(), 0, 2, d2, logl_logp + (size_t)c * steps * 2);
		}
	}
	if (rc == 0) {
		const long long dn[1] = {n_chains};
		rc = gwat_b200_dump_write(d, "/MCMC_METADATA/CHAIN TEMPERATURES", 0, 1, dn, temperatures);
		std::vector<int> zeros(n_chains, 0);
		if (rc == 0) rc = gwat_b200_dump_write(d, "/MCMC_METADATA/SUGGESTED TRIM LENGTHS", 1, 1, dn, trim_lengths ? trim_lengths : zeros.data());
	}
	if (rc == 0 && ac_values && n_cold > 0) {
		const long long da[2] = {n_cold, dimension};
		rc = gwat_b200_dump_write(d, "/MCMC_METADATA/AC VALUES", 1, 2, da, ac_values);
	}
	const int rc2 = gwat_b200_dump_close(d);
	return rc ? rc : rc2;
}

// write_flat_thin_output (src/mcmc_io_util.cpp:555-642) with stored autocorrelation lengths: the number of rows is
// (int)(mean over cold chains of (length - trim) / mean over cold chains of the largest ac value) -- count_indep_samples, :521-552 --
// and chain i contributes its steps j >= trim_i with j % max_ac_i == 0 while rows are left.  positions[n_cold][steps][dimension],
// ac_values[n_cold][dimension]; *n_rows receives the row count.
int gwat_b200_write_flat_thin_output(const char *path, int n_cold, int dimension, long long steps, const double *positions, const int *trim_lengths,
                                     const int *ac_values, long long *n_rows)
{
	if (!path || n_cold <= 0 || dimension <= 0 || steps < 0 || !positions || !ac_values) return GWAT_B200_ERR_ARG;
	std::vector<int> max_acs(n_cold, 1);
	double mean_ac = 0, mean_pos = 0;
	for (int i = 0; i < n_cold; i++) {
		int m = 1;
		for (int j = 0; j < dimension; j++)
			if (ac_values[(size_t)i * dimension + j] > m) m = ac_values[(size_t)i * dimension + j];
		max_acs[i] = m;
		mean_ac += m;
		mean_pos += (double)(steps - (trim_lengths ? trim_lengths[i] : 0));
	}
	mean_ac /= n_cold;
	mean_pos /= n_cold;
	const int indep = (int)(mean_pos / mean_ac);
	std::vector<double> flat((size_t)(indep > 0 ? indep : 0) * dimension);
	int ct = 0;
	for (int i = 0; i < n_cold; i++) {
		const long long begin = trim_lengths ? trim_lengths[i] : 0;
		for (long long j = begin; j < steps; j++)
			if (j % max_acs[i] == 0 && ct < indep) {
				std::memcpy(&flat[(size_t)ct * dimension], positions + ((size_t)i * steps + j) * dimension, sizeof(double) * dimension);
				ct++;
			}
	}
	gwat_b200_dump *d = nullptr;
	if (int rc = gwat_b200_dump_create(path, &d)) return rc;
	const long long dims[2] = {indep > 0 ? indep : 0, dimension};
	int rc = gwat_b200_dump_write(d, "/THINNED_MCMC_OUTPUT/THINNED FLATTENED CHAINS", 0, 2, dims, flat.data());
	const int rc2 = gwat_b200_dump_close(d);
	if (n_rows) *n_rows = dims[0];
	return rc ? rc : rc2;
}

}  // extern "C"
