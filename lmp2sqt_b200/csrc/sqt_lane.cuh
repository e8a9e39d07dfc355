// sqt_lane.cuh -- device helpers shared by the kernels that put a PARTICLE on each lane (sqt_mma.cu, sqt_vec.cu).
#pragma once

// cos(2 pi u) without SFU or libm: exact reduction in turns, fold to [0, 1/4], degree-8 polynomial in w^2 (Estrin form:
// depth 4, the callers are latency sensitive).  Same coefficients as cos2pi of sqt_kernels.cu.
__device__ __forceinline__ double mma_cos2pi(double u)
{
	const double MAGIC = 6755399441055744.0;
	double r = __dadd_rn(__dadd_rn(u, MAGIC), -MAGIC);
	double v = u - r;
	int hi = __double2hiint(v) & 0x7fffffff;
	double a = __hiloint2double(hi, __double2loint(v));
	const bool flip = hi > 0x3fd00000;
	double w = flip ? (0.5 - a) : a;
	double z = w * w;
	// degree-8 polynomial in z, Estrin grouping: depth 4 instead of Horner's 8 (this code sits in a latency-bound phase)
	const double z2 = z * z;
	const double q01 = fma(-19.739208802178705, z, 1.0);
	const double q23 = fma(-85.456817205973778, z, 64.939394022663905);
	const double q45 = fma(-26.426254049372801, z, 60.244641312691634);
	const double q67 = fma(-1.7132146974187872, z, 7.9034621397713947);
	const double z4 = z2 * z2;
	const double r03 = fma(q23, z2, q01);
	const double r47 = fma(q67, z2, q45);
	const double r07 = fma(r47, z4, r03);
	const double p = fma(0.27193002055395249 * z4, z4, r07);
	int ph = __double2hiint(p) ^ (flip ? 0x80000000 : 0);
	return __hiloint2double(ph, __double2loint(p));
}


// ring phase of origin slot nb at processed-frame index p (SURVEY.md section 8a closed form): false while idle, else lag tau
__device__ __forceinline__ bool mma_slot_state(int nb, int p, int T, int B, int *tau)
{
	int s_nb = (nb * T) / B;
	if (p < s_nb) return false;
	*tau = (p - s_nb) % T;
	return true;
}

// van Hove bin floor(sqrt(dr2)/rVal) (lmp2sqt.c:1429), bit-exact: an fp32 estimate (absolute error < 5e-5 bins up to nCS)
// decides alone unless it lies within 1e-3 of an integer; only then the exact comparison against the host's thresholds
// (bisected with the reference's own IEEE sqrt and divide) is made.  Returns nCS for "beyond the table".
__device__ __forceinline__ int lane_vanhove_bin(const double dr2, const float inv_rVal, const int nCS, const long long *s_thr)
{
	const float rf = __fsqrt_rn((float)dr2) * inv_rVal;
	int bi = (int)rf;
	bi = bi < 0 ? 0 : (bi > nCS ? nCS : bi);
	const float fr = rf - floorf(rf);
	if (!(fr >= 1e-3f && fr <= 0.999f)) {
		const long long xb = __double_as_longlong(dr2);
		if (xb >= s_thr[bi + 1]) ++bi;
		else if (xb < s_thr[bi]) --bi;
	}
	return bi;
}

// sum of n values p[0], p[stride], ... in a fixed order, four independent chains so that the loads overlap
// (the reduce kernels are latency bound: one dependent L2 access per addend otherwise)
__device__ __forceinline__ double strided_sum(const double *p, const long long stride, const int first, const int n, const int step)
{
	double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
	int k = first;
	for (; k + 3 * step < n; k += 4 * step) {
		v0 += p[(long long)k * stride]; v1 += p[(long long)(k + step) * stride];
		v2 += p[(long long)(k + 2 * step) * stride]; v3 += p[(long long)(k + 3 * step) * stride];
	}
	for (; k < n; k += step) v0 += p[(long long)k * stride];
	return (v0 + v1) + (v2 + v3);
}
