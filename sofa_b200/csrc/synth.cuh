// Synthetic trace text generated on the device (bench / test infrastructure, not part of
// the reference's path).  Byte-for-byte the same spec as sofa_b200/synth.py: every field
// is a pure function of (seed, stream, row, field) through a splitmix64 finaliser, so
// 100 M-event inputs can be produced directly in HBM and any row range can be
// regenerated on the host for the CPU baseline.
#pragma once
#if defined(__CUDACC__)
#include "kernels_emit.cuh"
#else
#include "rows.cuh"
#endif

namespace sofa {

SOFA_HD uint64_t mix64(uint64_t x) {
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
SOFA_HD uint64_t synth_rnd(uint64_t seed, uint64_t stream, uint64_t i, uint64_t f) {
    return mix64(seed + ((stream << 58) ^ ((i * 16 + f + 1) * 0x9E3779B97F4A7C15ull)));
}

template <class K> SOFA_HD void s_lit(K& k, const char* s) { for (int i = 0; s[i]; i++) k.put(s[i]); }
template <class K> SOFA_HD void s_u64(K& k, uint64_t v) {
    char b[24]; int n = fmt_u64(v, (char*)b);
    for (int i = 0; i < n; i++) k.put(b[i]);
}
template <class K> SOFA_HD void s_u64_w(K& k, uint64_t v, int width, bool left) {       // %Nd / %-Nd
    int n = ndigits_u64(v);
    if (!left) for (int i = n; i < width; i++) k.put(' ');
    s_u64(k, v);
    if (left) for (int i = n; i < width; i++) k.put(' ');
}
template <class K> SOFA_HD void s_u64_0(K& k, uint64_t v, int width) {                  // %0Nd
    char b[24]; fmt_u64_pad(v, width, (char*)b);
    for (int i = 0; i < width; i++) k.put(b[i]);
}
template <class K> SOFA_HD void s_hex(K& k, uint64_t v, int width) {                     // %Nx (space padded)
    char b[16]; int n = 0;
    if (v == 0) b[n++] = '0';
    while (v) { unsigned d = (unsigned)(v & 15); b[n++] = (char)(d < 10 ? '0' + d : 'a' + d - 10); v >>= 4; }
    for (int i = n; i < width; i++) k.put(' ');
    for (int i = n - 1; i >= 0; i--) k.put(b[i]);
}

SOFA_HD const char* synth_word(int i) {
    switch (i) {
        case 0: return "do_syscall_64"; case 1: return "schedule"; case 2: return "page_fault";
        case 3: return "tcp_sendmsg"; case 4: return "PyEval_EvalFrameDefault"; case 5: return "cudaLaunchKernel";
        case 6: return "ioctl"; case 7: return "futex_wait"; case 8: return "copy_user_enhanced_fast_string";
        case 9: return "__libc_start_main"; case 10: return "pthread_mutex_lock"; default: return "vfs_read";
    }
}
SOFA_HD const char* synth_lib(int i) {
    switch (i) {
        case 0: return "c"; case 1: return "pthread"; case 2: return "cuda"; case 3: return "cudart";
        case 4: return "python3.6m"; case 5: return "tensorflow_framework"; case 6: return "stdc++"; default: return "nccl";
    }
}
SOFA_HD const char* synth_kernel(int i) {
    switch (i) {
        case 0: return "volta_sgemm_128x64_nn";
        case 1: return "volta_scudnn_128x64_relu_interior_nn_v1";
        case 2: return "maxwell_gcgemm_32x32_nt";
        case 3: return "ncclAllReduceKernel_sum_f32";
        case 4: return "void tensorflow::functor::FillPhiloxRandomKernelLaunch<float, int=4>(float*, long)";
        case 5: return "void cudnn::detail::implicit_convolve_sgemm<float, float, int=128, int=6>(int, int)";
        case 6: return "cudnn_fw_conv_kernel";
        case 7: return "cudnn_bw_data_kernel";
        case 8: return "void Eigen::internal::EigenMetaKernel<Eigen::TensorEvaluator>(float)";
        case 9: return "AllReduceKernel_ring_ll";
        case 10: return "gemv2T_kernel_val";
        default: return "relu_fw_elementwise";
    }
}

// returns true when the symbol carries a "+0x.." offset
template <class K> SOFA_HD bool synth_perf_symbol(K& k, uint64_t idx) {
    int t = (int)(idx % 16);
    if (t == 0) { s_lit(k, "[unknown]"); return false; }
    if (t == 1) { s_lit(k, "cpu_idle_poll_"); s_u64(k, idx); return true; }
    if (t == 2) {   // _ZN10tensorflow<len>Op<k>Kernel7ComputeEv
        int len = 2 + ndigits_u64(idx) + 6;
        s_lit(k, "_ZN10tensorflow"); s_u64(k, (uint64_t)len); s_lit(k, "Op"); s_u64(k, idx); s_lit(k, "Kernel7ComputeEv");
        return true;
    }
    if (t == 3) {   // _ZN5Eigen8internal<len>gemm_pack_<k>3runEv
        int len = 10 + ndigits_u64(idx);
        s_lit(k, "_ZN5Eigen8internal"); s_u64(k, (uint64_t)len); s_lit(k, "gemm_pack_"); s_u64(k, idx); s_lit(k, "3runEv");
        return true;
    }
    if (t == 4) { s_lit(k, "memcpy_"); s_u64(k, idx); s_lit(k, "@plt"); return true; }
    s_lit(k, synth_word((int)(idx % 12))); k.put('_'); s_u64(k, idx);
    return true;
}

template <class K> SOFA_HD void synth_perf_line(K& k, uint64_t seed, uint64_t i, bool cpu_col) {
    uint64_t r0 = synth_rnd(seed, 1, i, 0), r1 = synth_rnd(seed, 1, i, 1), r2 = synth_rnd(seed, 1, i, 2), r3 = synth_rnd(seed, 1, i, 3);
    uint64_t tid = 4242 + (((r0 & 63) * ((r0 >> 6) & 63)) >> 6);
    uint64_t us = 5000ull * 1000000ull + i * 17 + (r0 >> 12) % 17;
    uint64_t period = 1 + (r1 % 3000000);
    bool kernel_ip = (r2 % 10) < 3;
    uint64_t ip = kernel_ip ? (0xffffffff80000000ull | ((r2 >> 8) & 0x0fffffffull)) : (0x7f0000000000ull | ((r2 >> 8) & 0xffffffffffull));
    uint64_t r20 = r3 & 0xfffff;
    uint64_t idx = (r20 * r20 * r20) >> 48;
    s_u64_w(k, 4242, 5, false); k.put('/'); s_u64_w(k, tid, 5, true);
    if (cpu_col) { s_lit(k, " ["); s_u64_0(k, (r1 >> 32) & 63, 3); k.put(']'); }
    k.put(' '); s_u64_w(k, us / 1000000, 7, false); k.put('.'); s_u64_0(k, us % 1000000, 6); s_lit(k, ": ");
    s_u64_w(k, period, 10, false); s_lit(k, " cycles:  "); s_hex(k, ip, 16); k.put(' ');
    if (synth_perf_symbol(k, idx)) { s_lit(k, "+0x"); s_hex(k, (r3 >> 20) & 0xfff, 0); }
    k.put(' ');
    if (kernel_ip) s_lit(k, "([kernel.kallsyms])");
    else {
        uint64_t j = (r3 >> 32) & 63;
        s_lit(k, "(/usr/lib/x86_64-linux-gnu/lib"); s_lit(k, synth_lib((int)(j & 7))); s_lit(k, ".so."); s_u64(k, j); k.put(')');
    }
    k.put('\n');
}

// variant 0: units s / ms / MB, 19 columns; variant 1: ms / us / KB with Src/Dst Dev/Ctx columns
template <class K> SOFA_HD void synth_gpu_line(K& k, uint64_t seed, uint64_t i, int variant) {
    uint64_t r0 = synth_rnd(seed, 2, i, 0), r1 = synth_rnd(seed, 2, i, 1), r2 = synth_rnd(seed, 2, i, 2);
    uint64_t us = 350000 + i * 23 + (r0 >> 12) % 23;
    uint64_t dur_ns = 800 + (r1 % 4000) * (1 + ((r1 >> 16) & 255));
    if (variant == 0) {
        s_u64(k, us / 1000000); k.put('.'); s_u64_0(k, us % 1000000, 6); k.put(',');
        s_u64(k, dur_ns / 1000000); k.put('.'); s_u64_0(k, dur_ns % 1000000, 6);
    } else {
        s_u64(k, us / 1000); k.put('.'); s_u64_0(k, us % 1000, 3); k.put(',');
        s_u64(k, dur_ns / 1000); k.put('.'); s_u64_0(k, dur_ns % 1000, 3);
    }
    uint64_t ctx = 1 + ((r0 >> 4) & 7), stream = 7 + ((r0 >> 7) & 31);
    uint64_t kind = r2 % 100;
    bool p2pcols = variant == 1;
    if (kind < 85) {
        uint64_t kk = (r2 >> 8) % 512;
        k.put(','); s_u64(k, 1 + ((r2 >> 20) & 1023)); s_lit(k, ",1,1,"); s_u64(k, 32ull << ((r2 >> 30) & 3)); s_lit(k, ",1,1,");
        s_u64(k, 16 + ((r2 >> 32) & 63)); s_lit(k, ",0,0,,,,,\"Tesla V100-SXM2-16GB ("); s_u64(k, ctx - 1); s_lit(k, ")\",\"");
        s_u64(k, ctx); s_lit(k, "\",\""); s_u64(k, stream); s_lit(k, "\",");
        if (p2pcols) s_lit(k, ",,,,");
        k.put('"'); s_lit(k, synth_kernel((int)(kk % 12)));
        if (kk >= 12) { s_lit(k, "_v"); s_u64(k, kk); }
        s_lit(k, "\"\n");
        return;
    }
    uint64_t nbytes = 4ull << ((r2 >> 8) % 24);
    s_lit(k, ",,,,,,,,,,");
    if (variant == 0) { s_u64(k, nbytes >> 20); k.put('.'); s_u64_0(k, ((nbytes & 0xfffff) * 1000000) >> 20, 6); }
    else { s_u64(k, nbytes >> 10); k.put('.'); s_u64_0(k, ((nbytes & 0x3ff) * 1000000) >> 10, 6); }
    uint64_t thr = 1000 + (r2 >> 40) % 24000000;
    k.put(','); s_u64(k, thr / 1000000); k.put('.'); s_u64_0(k, thr % 1000000, 6); k.put(',');
    const char* nm; const char* sm; const char* dm;
    bool is_p2p = false;
    if (kind < 91) { nm = "[CUDA memcpy HtoD]"; sm = "Pageable"; dm = "Device"; }
    else if (kind < 96) { nm = "[CUDA memcpy DtoH]"; sm = "Device"; dm = "Pageable"; }
    else if (kind < 98) { nm = "[CUDA memcpy DtoD]"; sm = "Device"; dm = "Device"; }
    else { nm = "[CUDA memcpy PtoP]"; sm = "Device"; dm = "Device"; is_p2p = true; }
    k.put('"'); s_lit(k, sm); s_lit(k, "\",\""); s_lit(k, dm); s_lit(k, "\",\"Tesla V100-SXM2-16GB ("); s_u64(k, ctx - 1);
    s_lit(k, ")\",\""); s_u64(k, ctx); s_lit(k, "\",\""); s_u64(k, stream); s_lit(k, "\",");
    if (p2pcols) {
        if (is_p2p) {
            uint64_t sc = 1 + ((r2 >> 16) & 7), dc = 1 + ((r2 >> 19) & 7);
            s_lit(k, "\"dev ("); s_u64(k, sc - 1); s_lit(k, ")\",\""); s_u64(k, sc); s_lit(k, "\",\"dev ("); s_u64(k, dc - 1);
            s_lit(k, ")\",\""); s_u64(k, dc); s_lit(k, "\",");
        } else s_lit(k, ",,,,");
    }
    k.put('"'); s_lit(k, nm); s_lit(k, "\"\n");
}

template <class K> SOFA_HD void synth_net_line(K& k, uint64_t seed, uint64_t i, int non_ip_pct) {
    uint64_t r0 = synth_rnd(seed, 3, i, 0), r1 = synth_rnd(seed, 3, i, 1);
    uint64_t us = 1600000000ull * 1000000ull + 250000 + i * 211 + (r0 >> 12) % 211;
    s_u64(k, us / 1000000); k.put('.'); s_u64_0(k, us % 1000000, 6);
    uint64_t sel = r1 % 100;
    if (sel < (uint64_t)non_ip_pct) {
        if ((r1 >> 8) & 1) {
            s_lit(k, " ARP, Request who-has 10.0.0."); s_u64(k, 1 + ((r1 >> 16) & 15)); s_lit(k, " tell 10.0.0.");
            s_u64(k, 1 + ((r1 >> 20) & 15)); s_lit(k, ", length 28\n");
        } else {
            s_lit(k, " IP 10.0.0."); s_u64(k, 1 + ((r1 >> 16) & 15)); s_lit(k, " > 10.0.0."); s_u64(k, 1 + ((r1 >> 20) & 15));
            s_lit(k, ": ICMP echo request, id 7, seq "); s_u64(k, (r1 >> 24) & 1023); s_lit(k, ", length 64\n");
        }
        return;
    }
    s_lit(k, " IP 10.0."); s_u64(k, (r0 >> 4) & 1); k.put('.'); s_u64(k, 1 + ((r0 >> 5) & 15)); k.put('.');
    s_u64(k, 1024 + ((r1 >> 8) & 0x7fff));
    s_lit(k, " > 10.0."); s_u64(k, (r0 >> 9) & 1); k.put('.'); s_u64(k, 1 + ((r0 >> 10) & 15)); k.put('.');
    s_u64(k, 1024 + ((r1 >> 24) & 0x7fff));
    if ((r1 >> 40) % 4 == 0) { s_lit(k, ": UDP, length "); s_u64(k, (r1 >> 44) % 1473); }
    else { s_lit(k, ": tcp "); s_u64(k, (r1 >> 44) % 65536); }
    k.put('\n');
}

template <class K> SOFA_HD void synth_line(K& k, int src, uint64_t seed, uint64_t i, int variant) {
    if (src == SRC_PERF) synth_perf_line(k, seed, i, variant != 0);
    else if (src == SRC_GPU) synth_gpu_line(k, seed, i, variant);
    else synth_net_line(k, seed, i, variant);
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(EMIT_TPB) k_synth_size(int src, uint64_t seed, uint64_t row0, uint64_t n, int variant, uint32_t* rowlen,
                                                         uint64_t* blockoff) {
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * EMIT_TPB + threadIdx.x;
    uint32_t len = 0;
    if (i < n) { LenSink l; synth_line(l, src, seed, row0 + i, variant); len = (uint32_t)l.len; rowlen[i] = len; }
    uint32_t total;
    block_excl_scan(len, &total, sm);
    if (threadIdx.x == 0) blockoff[blockIdx.x] = total;
}
__global__ void __launch_bounds__(EMIT_TPB) k_synth_write(int src, uint64_t seed, uint64_t row0, uint64_t n, int variant, const uint32_t* rowlen,
                                                          const uint64_t* blockoff, char* out, uint64_t base) {
    __shared__ __align__(16) unsigned char stage[EMIT_STAGE_BYTES + 32];
    __shared__ uint32_t sm[33];
    uint64_t i = (uint64_t)blockIdx.x * EMIT_TPB + threadIdx.x;
    uint32_t len = i < n ? rowlen[i] : 0;
    uint32_t total;
    uint32_t ex = block_excl_scan(len, &total, sm);
    uint64_t tile_off = base + blockoff[blockIdx.x];
    bool staged = total <= EMIT_STAGE_BYTES;
    if (len) {
        char* dst = staged ? reinterpret_cast<char*>(stage) + (tile_off & 15) + ex : out + tile_off + ex;
        RawWrite<char*> w{dst, 0};
        synth_line(w, src, seed, row0 + i, variant);
    }
    __syncthreads();
    if (staged && total) flush_staged(stage, out, tile_off, total);
}
#endif

}  // namespace sofa
