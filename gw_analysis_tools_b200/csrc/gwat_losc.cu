// LOSC (GWOSC) text files -> frequency-domain strain, PSDs and the frequency grid of an analysis.
//
// Reference: allocate_LOSC_data (src/io_util.cpp:523-661) with read_LOSC_data_file (:403-464), read_LOSC_PSD_file (:466-500),
// tukey_window (src/util.cpp:1736-1757) and an FFTW forward transform per detector (src/util.cpp:964-969).  It reads one
// strain file per detector (three header lines: sampling rate on the second, GPS start and duration on the third, then the
// samples) and one PSD file (a header line, then rows "f S_1 ... S_D"), cuts T_obs = 1/df seconds of strain ending
// `post_merger_duration` after the trigger, applies a Tukey window with alpha = 0.8/T_obs, transforms, multiplies by dt and
// keeps the bins of the PSD's frequency range.
// Here the files are parsed on the host (they are text), the window, the transform (one batched cuFFT Z2Z, as the reference
// plans FFTW_FORWARD on a complex array with zero imaginary part) and the scaling/cut run on the GPU.
#include <cuda_runtime.h>
#include <cufft.h>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "gwat_engine_internal.h"

namespace {

// window[l] * x[l] as the real part of a complex array, zero-padded to n (tukey_window, src/util.cpp:1736-1757)
__global__ void k_tukey_apply(const double *__restrict__ x, int n_sel, int n, double alpha, cufftDoubleComplex *__restrict__ out)
{
	const int l = blockIdx.x * blockDim.x + threadIdx.x;
	const int d = blockIdx.y;
	if (l >= n) return;
	const double imin = (double)(int)(alpha * (double)(n - 1) / 2.0);
	const double imax = (double)(int)((double)(n - 1) * (1. - alpha / 2.));
	double w;
	if (l < imin) w = 0.5 * (1 + cos(M_PI * (((double)l) / imin - 1)));
	else if (l < imax) w = 1;
	else w = 0.5 * (1 + cos(M_PI * (((double)l) / imin - 2. / alpha + 1.)));
	const double v = l < n_sel ? x[(size_t)d * n_sel + l] * w : 0.0;
	out[(size_t)d * n + l] = cufftDoubleComplex{v, 0.0};
}

// data[d][l] = fft[d][i0 + l] * dt      (src/io_util.cpp:637-647)
__global__ void k_cut_scale(const cufftDoubleComplex *__restrict__ fft, int n, int i0, int count, double dt, double *__restrict__ re,
                            double *__restrict__ im)
{
	const int l = blockIdx.x * blockDim.x + threadIdx.x;
	const int d = blockIdx.y;
	if (l >= count) return;
	const cufftDoubleComplex v = fft[(size_t)d * n + i0 + l];
	re[(size_t)d * count + l] = v.x * dt;
	im[(size_t)d * count + l] = v.y * dt;
}

struct StrainFile {
	double fs = 0, start = 0, duration = 0;
	std::vector<double> x;
};

// read_LOSC_data_file, src/io_util.cpp:403-464
bool read_strain(const char *path, StrainFile &f)
{
	std::ifstream in(path);
	if (!in) return false;
	std::string line;
	int j = 0;
	while (std::getline(in, line)) {
		if (j > 2) {
			std::stringstream ls(line);
			std::string item;
			while (std::getline(ls, item, ',')) {
				char *end = nullptr;
				const double v = std::strtod(item.c_str(), &end);
				if (end != item.c_str()) f.x.push_back(v);
			}
		} else if (j == 1) {
			std::istringstream iss(line);
			for (std::string s; iss >> s;)
				if (std::isdigit((unsigned char)s[0])) f.fs = std::strtod(s.c_str(), nullptr);
		} else if (j == 2) {
			std::istringstream iss(line);
			int k = 0;
			for (std::string s; iss >> s;) {
				if (!std::isdigit((unsigned char)s[0])) continue;
				if (k == 0) {
					f.start = std::strtod(s.c_str(), nullptr);
					k++;
				} else f.duration = std::strtod(s.c_str(), nullptr);
			}
		}
		j++;
	}
	return f.fs > 0 && f.duration > 0 && !f.x.empty();
}

// read_LOSC_PSD_file, src/io_util.cpp:466-500: header row, then whitespace-separated rows of 1 + D numbers
bool read_psd(const char *path, int cols, std::vector<double> &flat)
{
	std::ifstream in(path);
	if (!in) return false;
	std::string line;
	bool header = true;
	while (std::getline(in, line)) {
		if (header) {
			header = false;
			continue;
		}
		std::istringstream iss(line);
		for (std::string s; iss >> s;) flat.push_back(std::strtod(s.c_str(), nullptr));
	}
	return !flat.empty() && flat.size() % cols == 0 && flat.size() / cols >= 2;
}

}  // namespace

extern "C" int gwat_b200_losc_prepare(gwat_b200_ctx *ctx, int num_detectors, const char *const *data_files, const char *psd_file,
                                      double trigger_time, double post_merger_duration, int capacity, int *length, double *frequencies,
                                      double *psd, double *data_re, double *data_im)
{
	using gwat_internal::set_error;
	if (!ctx) return GWAT_B200_ERR_ARG;
	if (num_detectors < 1 || num_detectors > GWAT_B200_MAX_DETECTORS || !data_files || !psd_file || !length)
		return set_error(ctx, GWAT_B200_ERR_ARG, "losc_prepare: bad arguments");
	const int D = num_detectors;
	std::vector<double> ptab;
	if (!read_psd(psd_file, D + 1, ptab)) return set_error(ctx, GWAT_B200_ERR_STATE, std::string("losc_prepare: cannot read PSD file ") + psd_file);
	const int rows = (int)(ptab.size() / (D + 1));
	*length = rows;
	if (capacity < rows || !frequencies || !psd || !data_re || !data_im)
		return set_error(ctx, GWAT_B200_ERR_ARG, "losc_prepare: output arrays too small (needed length returned)");
	std::vector<StrainFile> files(D);
	for (int d = 0; d < D; d++) {
		if (!data_files[d] || !read_strain(data_files[d], files[d]))
			return set_error(ctx, GWAT_B200_ERR_STATE, std::string("losc_prepare: cannot read strain file ") + (data_files[d] ? data_files[d] : "(null)"));
	}
	// the reference takes fs, start and duration from the LAST file it reads and assumes the others agree
	const double fs = files[D - 1].fs, duration = files[D - 1].duration, file_start = files[D - 1].start;
	for (int d = 0; d < D; d++)
		if (files[d].fs != fs || files[d].duration != duration || files[d].start != file_start)
			return set_error(ctx, GWAT_B200_ERR_ARG, "losc_prepare: strain files differ in sampling rate, start or duration");
	for (int j = 0; j < rows; j++) {
		frequencies[j] = ptab[(size_t)j * (D + 1)];
		for (int d = 0; d < D; d++) psd[(size_t)d * rows + j] = ptab[(size_t)j * (D + 1) + d + 1];
	}
	const double Tobs = 1. / (frequencies[1] - frequencies[0]);
	const double df = 1. / Tobs, dt = 1. / fs;
	const int n_trim = (int)(Tobs * fs);
	const int N = (int)(fs * duration);
	const double time_start = trigger_time - (Tobs - post_merger_duration), time_end = trigger_time + post_merger_duration;
	if (time_start < file_start || time_end > file_start + duration)
		return set_error(ctx, GWAT_B200_ERR_ARG, "losc_prepare: trigger time does not fit inside the data file");
	if (n_trim < 2) return set_error(ctx, GWAT_B200_ERR_ARG, "losc_prepare: observation time shorter than two samples");
	// samples with time_start < t_i <= time_end, t_i = file_start + i dt  (src/io_util.cpp:603-611)
	int i_first = -1, n_sel = 0;
	for (int i = 0; i < N; i++) {
		const double t = file_start + i * dt;
		if (t > time_start && t <= time_end) {
			if (i_first < 0) i_first = i;
			n_sel++;
		}
	}
	if (n_sel > n_trim) n_sel = n_trim;
	for (int d = 0; d < D; d++)
		if (i_first < 0 || (size_t)(i_first + n_sel) > files[d].x.size())
			return set_error(ctx, GWAT_B200_ERR_STATE, "losc_prepare: strain file shorter than its header says");
	// bins with fmin <= i df <= fmax  (src/io_util.cpp:629-647)
	const double fmin = frequencies[0], fmax = frequencies[rows - 1];
	int i0 = -1, count = 0;
	for (int i = 0; i < n_trim; i++) {
		const double f = i * df;
		if (f >= fmin && f <= fmax) {
			if (i0 < 0) i0 = i;
			count++;
		}
	}
	if (count > rows) count = rows;
	if (count <= 0) return set_error(ctx, GWAT_B200_ERR_ARG, "losc_prepare: the PSD's frequency range is outside the transform's");

	std::lock_guard<std::mutex> lock(ctx->mu);
	if (cudaSetDevice(ctx->device) != cudaSuccess) return set_error(ctx, GWAT_B200_ERR_CUDA, "cudaSetDevice failed");
	cudaStream_t st = ctx->stream;
	double *d_x = nullptr, *d_out = nullptr;
	cufftDoubleComplex *d_a = nullptr, *d_b = nullptr;
	cufftHandle plan = 0;
	bool have_plan = false;
	int rc = GWAT_B200_OK;
	auto fail = [&](const char *what) { rc = set_error(ctx, GWAT_B200_ERR_CUDA, what); };
	do {
		if (cudaMalloc((void **)&d_x, sizeof(double) * (size_t)D * n_sel) != cudaSuccess ||
		    cudaMalloc((void **)&d_a, sizeof(cufftDoubleComplex) * (size_t)D * n_trim) != cudaSuccess ||
		    cudaMalloc((void **)&d_b, sizeof(cufftDoubleComplex) * (size_t)D * n_trim) != cudaSuccess ||
		    cudaMalloc((void **)&d_out, sizeof(double) * (size_t)2 * D * count) != cudaSuccess) {
			fail("losc_prepare: cudaMalloc failed");
			break;
		}
		for (int d = 0; d < D; d++)
			cudaMemcpyAsync(d_x + (size_t)d * n_sel, files[d].x.data() + i_first, sizeof(double) * n_sel, cudaMemcpyHostToDevice, st);
		const double alpha = 2 * .4 / Tobs;
		k_tukey_apply<<<dim3((n_trim + 255) / 256, D), 256, 0, st>>>(d_x, n_sel, n_trim, alpha, d_a);
		if (cufftPlan1d(&plan, n_trim, CUFFT_Z2Z, D) != CUFFT_SUCCESS) {
			fail("losc_prepare: cufftPlan1d failed");
			break;
		}
		have_plan = true;
		cufftSetStream(plan, st);
		if (cufftExecZ2Z(plan, d_a, d_b, CUFFT_FORWARD) != CUFFT_SUCCESS) {
			fail("losc_prepare: cufftExecZ2Z failed");
			break;
		}
		k_cut_scale<<<dim3((count + 255) / 256, D), 256, 0, st>>>(d_b, n_trim, i0, count, dt, d_out, d_out + (size_t)D * count);
		ctx->launches += 3;
		std::vector<double> host((size_t)2 * D * count);
		if (cudaMemcpyAsync(host.data(), d_out, sizeof(double) * host.size(), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
		    cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) {
			fail("losc_prepare: kernel or copy failed");
			break;
		}
		for (int d = 0; d < D; d++)
			for (int l = 0; l < rows; l++) {
				data_re[(size_t)d * rows + l] = l < count ? host[(size_t)d * count + l] : 0.0;
				data_im[(size_t)d * rows + l] = l < count ? host[(size_t)(D + d) * count + l] : 0.0;
			}
	} while (false);
	if (have_plan) cufftDestroy(plan);
	cudaFree(d_x);
	cudaFree(d_a);
	cudaFree(d_b);
	cudaFree(d_out);
	return rc;
}
