/* A plain-C caller of libs1s2_b200.so: no Python, no PyTorch -- the binding a non-Python host of the reference's path
 * would write (include/s1s2_b200.h).  Builds the network with constant synthetic weights, runs one model call
 * (s1s2_forward) and one two-step v-DDIM chain through the host-buffer entry (s1s2_sample_host), prints checksums.
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include tests/cabi/host_example.c -o host_example \
 *       -L <dir of libs1s2_b200.so> -ls1s2_b200 -L /usr/local/cuda/lib64 -lcudart -lm
 *
 * Exit code 0 = every call succeeded and the outputs are finite; 77 = no usable GPU (the library has no CPU fallback). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "s1s2_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != S1S2_OK) {                                                    \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, h ? s1s2_last_error(h) : s1s2_global_error()); \
            return 1;                                                            \
        }                                                                        \
    } while (0)

/* the 34 tensors of UNetSmall(8, 4, base_ch) in state_dict order: name and element count */
static int param_list(int b, const char** names, long long* numel) {
    static char buf[34][32];
    int n = 0, c = b, i;
    const char* enc[3] = {"down1", "down2", "down3"};
    const char* up[3] = {"up3", "up2", "up1"};
    const char* dec[3] = {"conv3", "conv2", "conv1"};
#define ADD(fmt, a, cnt) do { snprintf(buf[n], sizeof(buf[n]), fmt, a); names[n] = buf[n]; numel[n] = (cnt); ++n; } while (0)
    ADD("inc.0.%s", "weight", (long long)b * 9 * 9); ADD("inc.0.%s", "bias", b);
    for (i = 0; i < 3; ++i) {
        ADD("%s.0.0.weight", enc[i], (long long)2 * c * c * 9); ADD("%s.0.0.bias", enc[i], 2 * c);
        ADD("%s.0.2.weight", enc[i], (long long)2 * c * 2 * c * 9); ADD("%s.0.2.bias", enc[i], 2 * c);
        c *= 2;
    }
    for (i = 0; i < 3; ++i) {
        ADD("%s.weight", up[i], (long long)c * (c / 2) * 4); ADD("%s.bias", up[i], c / 2);
        ADD("%s.0.weight", dec[i], (long long)(c / 2) * c * 9); ADD("%s.0.bias", dec[i], c / 2);
        ADD("%s.2.weight", dec[i], (long long)(c / 2) * (c / 2) * 9); ADD("%s.2.bias", dec[i], c / 2);
        c /= 2;
    }
    ADD("outc.%s", "weight", 4LL * b); ADD("outc.%s", "bias", 4);
#undef ADD
    return n;
}

int main(void) {
    s1s2_handle* h = NULL;
    const int B = 2, H = 32, W = 32, base_ch = 96;
    const size_t img = (size_t)B * 4 * H * W;
    const char* names[34];
    long long numel[34];
    const float* ptrs[34];
    int ndev = 0, n, i;
    size_t k;

    printf("s1s2 abi %d\n", s1s2_abi_version());
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        int rc = s1s2_create(&h, 0, 8, 4, base_ch, H, W, B);
        printf("no CUDA device: s1s2_create -> %d (%s)\n", rc, s1s2_global_error());
        return rc != S1S2_OK ? 77 : 1;              /* must fail loudly: there is no CPU fallback */
    }
    CHECK(s1s2_create(&h, 0, 8, 4, base_ch, H, W, B));
    n = param_list(base_ch, names, numel);
    for (i = 0; i < n; ++i) {                        /* small deterministic weights, uploaded as float32 like a state_dict */
        float* host = (float*)malloc(sizeof(float) * (size_t)numel[i]);
        float* dev = NULL;
        for (k = 0; k < (size_t)numel[i]; ++k) host[k] = 0.02f * sinf(0.37f * (float)k + (float)i);
        if (cudaMalloc((void**)&dev, sizeof(float) * (size_t)numel[i]) != cudaSuccess) return 1;
        cudaMemcpy(dev, host, sizeof(float) * (size_t)numel[i], cudaMemcpyHostToDevice);
        free(host);
        ptrs[i] = dev;
    }
    CHECK(s1s2_load_weights(h, n, names, ptrs, (const int64_t*)numel, NULL));

    {   /* model(torch.cat([x_t, x_cond], 1), t_idx) */
        float *x = NULL, *out = NULL, *host = (float*)malloc(sizeof(float) * 2 * img);
        int64_t t_host[2] = {999, 20}, *t = NULL;
        double sum = 0.0;
        for (k = 0; k < 2 * img; ++k) host[k] = cosf(0.01f * (float)k);
        cudaMalloc((void**)&x, sizeof(float) * 2 * img);
        cudaMalloc((void**)&out, sizeof(float) * img);
        cudaMalloc((void**)&t, sizeof(t_host));
        cudaMemcpy(x, host, sizeof(float) * 2 * img, cudaMemcpyHostToDevice);
        cudaMemcpy(t, t_host, sizeof(t_host), cudaMemcpyHostToDevice);
        CHECK(s1s2_forward(h, x, t, out, B, NULL));
        cudaMemcpy(host, out, sizeof(float) * img, cudaMemcpyDeviceToHost);
        for (k = 0; k < img; ++k) { if (!isfinite(host[k])) { fprintf(stderr, "non-finite output\n"); return 1; } sum += host[k]; }
        printf("forward: sum %.6f\n", sum);
        free(host);
    }
    {   /* two v-DDIM steps (t = 999 -> 0) through the host-buffer entry */
        float *cond = (float*)malloc(sizeof(float) * img), *noise = (float*)malloc(sizeof(float) * img), *res = (float*)malloc(sizeof(float) * img);
        s1s2_step st[2];
        double sum = 0.0;
        memset(st, 0, sizeof(st));
        st[0].t = 999; st[0].kind = S1S2_STEP_V_DDIM; st[0].noise_index = -1;
        st[0].c0 = 4.93e-5f; st[0].c1 = 1.0f; st[0].c2 = 0.99998f; st[0].c3 = 0.00643f;      /* sqrt(abar_999), sqrt(1-abar_999), sqrt(abar_0), sqrt(1-abar_0) */
        st[1].t = 0; st[1].kind = S1S2_STEP_V_DDIM; st[1].flags = S1S2_STEP_FINAL; st[1].noise_index = -1;
        st[1].c0 = 0.99998f; st[1].c1 = 0.00643f;
        for (k = 0; k < img; ++k) { cond[k] = sinf(0.02f * (float)k); noise[k] = cosf(0.03f * (float)k); }
        CHECK(s1s2_sample_host(h, st, 2, cond, noise, 1.0f, res, B, NULL));
        for (k = 0; k < img; ++k) { if (!(res[k] >= 0.f && res[k] <= 1.f)) { fprintf(stderr, "result outside [0, 1]\n"); return 1; } sum += res[k]; }
        printf("sample_host: sum %.6f, %lld kernel launches\n", sum, (long long)s1s2_launch_count(h));
        free(cond); free(noise); free(res);
    }
    s1s2_destroy(h);
    printf("ok\n");
    return 0;
}
