// CPU emulation of the fused 8192-point convolution stages (csrc/conv8k.cuh) against a direct float64 circular convolution,
// and of their agreement with the transform order of fft8k_dif (the order zc_spectrum8k_kernel's float64 FFT produces).
// Built and run by tests/test_host_conv8k.py with g++ (no GPU, no CUDA toolkit needed).
#include "common.cuh"
#include "fft4096.cuh"
#include "conv8k.cuh"
#include <complex>
#include <cstdio>
#include <random>
#include <vector>
using namespace ofs;
typedef std::complex<double> cd;

int main()
{
    const double PI = 3.14159265358979323846;
    // twiddle tables as zc_twiddle_kernel / zc_twiddle8_kernel write them
    std::vector<unsigned char> twbuf((ZF / 2) * (sizeof(double2) + sizeof(float2)));
    double2 *tw = reinterpret_cast<double2 *>(twbuf.data());
    float2 *twf = reinterpret_cast<float2 *>(tw + ZF / 2);
    for (int i = 0; i < ZF / 2; ++i) {
        const double c = cos(-2 * PI * i / ZF), s = sin(-2 * PI * i / ZF);
        tw[i] = make_double2(c, s); twf[i] = make_float2((float)c, (float)s);
    }
    std::vector<float2> tw8f(256); std::vector<double2> tw8d(256);
    for (int i = 0; i < 256; ++i) {
        const double c = cos(-2 * PI * i / ZF8), s = sin(-2 * PI * i / ZF8);
        tw8d[i] = make_double2(c, s); tw8f[i] = make_float2((float)c, (float)s);
    }
    std::mt19937 rng(7);
    std::normal_distribution<double> nd;
    const int nr = 2048;
    std::vector<cd> x(ZF8), g(ZF8, cd(0, 0));
    for (auto &v : x) v = cd(nd(rng), nd(rng));
    for (int m = 0; m < nr; ++m) g[m] = cd(nd(rng), nd(rng));

    // ---- filter spectrum in transform order through the float64 fft8k_dif, emulated pass by pass (one loop per barrier)
    std::vector<double2> ad(ZFP8);
    for (int m = 0; m < ZF8; ++m) ad[zpad8(m)] = make_double2(g[m].real(), g[m].imag());
    for (int t = 0; t < 256; ++t)                       // radix-2 stage of fft8k_dif
        for (int q = 0; q < 16; ++q) {
            const int n = zpad(t + 256 * q);
            const double2 u = ad[n], v = ad[ZFP + n];
            ad[n] = cadd(u, v);
            const double2 d = csub(u, v);
            ad[ZFP + n] = q == 0 ? cmul(d, tw8d[t]) : cmul(cmul(d, tw8d[t]), w32_const<double2>(q));
        }
    for (int h = 0; h < 2; ++h) {                       // fft_dif, three passes
        double2 *a = ad.data() + h * ZFP;
        for (int t = 0; t < 256; ++t) {
            double2 v[16], w[16];
            for (int q = 0; q < 16; ++q) v[q] = a[zpad(q * 256 + t)];
            dft16<double2, false>(v);
            tw_powers<double2>(tw4096<double2>(tw, t), w);
            for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], w[q]);
            for (int q = 0; q < 16; ++q) a[zpad(q * 256 + t)] = v[q];
        }
        for (int t = 0; t < 256; ++t) {
            const int k0 = t >> 4, n0 = t & 15;
            double2 v[16], w[16];
            for (int q = 0; q < 16; ++q) v[q] = a[zpad(k0 * 256 + q * 16 + n0)];
            dft16<double2, false>(v);
            tw_powers<double2>(tw4096<double2>(tw, 16 * n0), w);
            for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], w[q]);
            for (int q = 0; q < 16; ++q) a[zpad(k0 * 256 + q * 16 + n0)] = v[q];
        }
        for (int t = 0; t < 256; ++t) {
            double2 v[16];
            for (int q = 0; q < 16; ++q) v[q] = a[t * 17 + q];
            dft16<double2, false>(v);
            for (int q = 0; q < 16; ++q) a[t * 17 + q] = v[q];
        }
    }
    std::vector<float2> Gp(ZF8);
    for (int m = 0; m < ZF8; ++m) {                      // the store of zc_spectrum8k_kernel
        const double2 gg = ad[zpad8(m)];
        const int h = m >> 12, t = (m & (ZF - 1)) >> 4, q = m & 15;
        Gp[conv8k_gidx(h, t, q)] = make_float2((float)(gg.x / ZF8), (float)(gg.y / ZF8));
    }
    // sanity: the transform-order array is a permutation of the true spectrum (compare multisets through a checksum of |G|^2)
    // ---- the fused pipeline, thread by thread
    std::vector<float2> a(ZFP8), out(ZF8);
    std::vector<float> en(ZF8);
    auto all_threads = [&](auto fn) { for (int t = 0; t < 256; ++t) { threadIdx.x = t; fn(); } };
    all_threads([&] {
        conv8k_stage_a(a.data(), conv8k_seeds_ae(tw), tw8f[threadIdx.x], [&](int m) { return make_float2((float)x[m].real(), (float)x[m].imag()); },
                       [&](int m, float e) { en[m] = e; });
    });
    all_threads([&] { conv8k_stage_b(a.data(), conv8k_seeds_bd(tw)); });
    all_threads([&] { conv8k_stage_c<false>(a.data(), Gp.data(), nullptr, nullptr); });
    all_threads([&] { conv8k_stage_d(a.data(), conv8k_seeds_bd(tw)); });
    all_threads([&] { conv8k_stage_e(a.data(), conv8k_seeds_ae(tw), tw8f[threadIdx.x], conv8k_no_pre(), [&](int m, float2 y, int) { out[m] = y; }); });

    // ---- reference: circular convolution of x (as float) with g, float64
    double maxerr = 0, scale = 0;
    std::vector<cd> xf(ZF8);
    for (int m = 0; m < ZF8; ++m) xf[m] = cd((float)x[m].real(), (float)x[m].imag());
    for (int k = 0; k < ZF8; k += 37) {                  // sampled outputs (a direct sum is 2048 terms each)
        cd s(0, 0);
        for (int m = 0; m < nr; ++m) s += g[m] * xf[(k - m + ZF8) % ZF8];
        maxerr = std::max(maxerr, std::abs(s - cd(out[k].x, out[k].y)));
        scale = std::max(scale, std::abs(s));
    }
    double een = 0;
    for (int m = 0; m < ZF8; ++m) een = std::max(een, std::abs((double)en[m] - std::norm(xf[m])));
    printf("conv8k: max |err| = %.3e of scale %.3e (rel %.3e), energy err %.3e\n", maxerr, scale, maxerr / scale, een);

    // ---- two-filter form: second spectrum = all-pass delay by 5 samples; the stash must hold x delayed
    std::vector<cd> g2(ZF8, cd(0, 0)); g2[5] = 1.0;
    for (int m = 0; m < ZF8; ++m) ad[zpad8(m)] = make_double2(g2[m].real(), g2[m].imag());
    // spectrum of a delay: W^(5 k); use the pipeline itself to get transform order: forward stages on g2 in float are accurate enough
    std::vector<float2> a2(ZFP8), Gp2(ZF8), stash(ZF8), out2(ZF8);
    all_threads([&] { conv8k_stage_a(a2.data(), conv8k_seeds_ae(tw), tw8f[threadIdx.x], [&](int m) { return make_float2((float)g2[m].real(), (float)g2[m].imag()); }, [](int, float) {}); });
    all_threads([&] { conv8k_stage_b(a2.data(), conv8k_seeds_bd(tw)); });
    for (int h = 0; h < 2; ++h)
        for (int t = 0; t < 256; ++t) {
            float2 v[16];
            for (int q = 0; q < 16; ++q) v[q] = a2[h * ZFP + t * 17 + q];
            pk::dft16<false>(v);
            for (int q = 0; q < 16; ++q) Gp2[conv8k_gidx(h, t, q)] = make_float2(v[q].x / ZF8, v[q].y / ZF8);
        }
    all_threads([&] { conv8k_stage_a(a.data(), conv8k_seeds_ae(tw), tw8f[threadIdx.x], [&](int m) { return make_float2((float)x[m].real(), (float)x[m].imag()); }, [](int, float) {}); });
    all_threads([&] { conv8k_stage_b(a.data(), conv8k_seeds_bd(tw)); });
    all_threads([&] { conv8k_stage_c<true>(a.data(), Gp.data(), Gp2.data(), stash.data()); });
    all_threads([&] { conv8k_stage_d(a.data(), conv8k_seeds_bd(tw)); });
    all_threads([&] { conv8k_stage_e(a.data(), conv8k_seeds_ae(tw), tw8f[threadIdx.x], conv8k_no_pre(), [&](int m, float2 y, int) { out[m] = y; }); });
    all_threads([&] { conv8k_unstash(a.data(), stash.data()); });
    all_threads([&] { conv8k_stage_d(a.data(), conv8k_seeds_bd(tw)); });
    all_threads([&] { conv8k_stage_e(a.data(), conv8k_seeds_ae(tw), tw8f[threadIdx.x], conv8k_no_pre(), [&](int m, float2 y, int) { out2[m] = y; }); });
    double e1 = 0, e2 = 0;
    for (int k = 0; k < ZF8; k += 37) {
        cd s(0, 0);
        for (int m = 0; m < nr; ++m) s += g[m] * xf[(k - m + ZF8) % ZF8];
        e1 = std::max(e1, std::abs(s - cd(out[k].x, out[k].y)));
        e2 = std::max(e2, std::abs(xf[(k - 5 + ZF8) % ZF8] - cd(out2[k].x, out2[k].y)));
    }
    printf("conv8k two filters: first rel err %.3e, delayed copy abs err %.3e\n", e1 / scale, e2);
    // ---- the array as 32 independent 256-point transforms (block-FFT Park kernel): stage B + stage C (forward half) on every
    //      256-block, position-wise product of two spectra, stage C (inverse half) + stage D: linear convolution of two 128-blocks
    std::vector<float2> ab(ZFP8, make_float2(0.f, 0.f));
    std::vector<cd> u0(128), u1(128);
    for (auto &v : u0) v = cd(nd(rng), nd(rng));
    for (auto &v : u1) v = cd(nd(rng), nd(rng));
    for (int e = 0; e < 128; ++e) {
        ab[conv8k_blk(3, e)] = make_float2((float)u0[e].real(), (float)u0[e].imag());      // slots 3 and 21: both halves of the array
        ab[conv8k_blk(21, e)] = make_float2((float)u1[e].real(), (float)u1[e].imag());
    }
    all_threads([&] { conv8k_stage_b(ab.data(), conv8k_seeds_bd(tw)); });
    all_threads([&] { conv8k_stage_c_fwd(ab.data()); });
    for (int f = 0; f < 256; ++f) ab[conv8k_blk(30, f)] = pk::mul(ab[conv8k_blk(3, f)], ab[conv8k_blk(21, f)]);
    all_threads([&] { conv8k_stage_c_inv(ab.data()); });
    all_threads([&] { conv8k_stage_d(ab.data(), conv8k_seeds_bd(tw)); });
    double e3 = 0, sc3 = 0;
    for (int k = 0; k < 255; ++k) {
        cd s(0, 0);
        for (int m = 0; m < 128; ++m)
            if (k - m >= 0 && k - m < 128) s += cd((float)u0[m].real(), (float)u0[m].imag()) * cd((float)u1[k - m].real(), (float)u1[k - m].imag());
        const float2 g = ab[conv8k_blk(30, k)];
        e3 = std::max(e3, std::abs(s - cd(g.x / 256.0, g.y / 256.0)));
        sc3 = std::max(sc3, std::abs(s));
    }
    printf("256-point blocks: linear convolution rel err %.3e\n", e3 / sc3);
    const bool ok = maxerr / scale < 2e-6 && een < 1e-5 && e1 / scale < 2e-6 && e2 < 2e-5 && e3 / sc3 < 2e-6;
    printf(ok ? "PASS\n" : "FAIL\n");
    return ok ? 0 : 1;
}
