// A host without Python: raw model arrays -> bf_model_create -> bf_workspace_bytes / bf_frames_bind -> bf_pack_keypoints /
// bf_init_theta -> bf_fit_run, through include/bodyfit_b200.h only.  tests/test_gpu_parity.py::test_c_host_program builds this
// file with nvcc, feeds it a dump of the synthetic model + inputs and compares the fitted parameters with the Python session's.
//
// Input file: int32 magic 0xB200C057; then, for each pointer member of BfModelDesc in declaration order, int64 nbytes (0 = NULL)
// followed by the bytes; the 16 int32 members of BfModelDesc; int32 B, Nv, K, ld_poses, n_iters, hand_face; then float32 arrays
// kp_raw [B,Nv,K,3], cams [Nv,12], poses [B,ld_poses], betas [B,10].  Output file: float32 theta [B,NP].
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "bodyfit_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define BF(x) do { int r_ = (x); if (r_) { fprintf(stderr, "%s failed (%d): %s\n", #x, r_, bf_last_error()); return 3; } } while (0)

static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: fit_host input.bin theta_out.bin\n"); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    int32_t magic = 0;
    if (!rd(f, &magic, 4) || (uint32_t)magic != 0xB200C057u) { fprintf(stderr, "bad input file\n"); return 1; }
    BfModelDesc d;
    memset(&d, 0, sizeof(d));
    const int n_ptr = (int)(offsetof(BfModelDesc, is_smplx) / sizeof(void*));
    std::vector<std::vector<char> > keep(n_ptr);
    for (int i = 0; i < n_ptr; ++i) {
        int64_t nb = 0;
        if (!rd(f, &nb, 8)) return 1;
        if (nb > 0) {
            keep[i].resize((size_t)nb);
            if (!rd(f, keep[i].data(), (size_t)nb)) return 1;
            ((const void**)&d)[i] = keep[i].data();
        }
    }
    if (!rd(f, &d.is_smplx, 16 * sizeof(int32_t))) return 1;
    int32_t hdr[6];
    if (!rd(f, hdr, sizeof(hdr))) return 1;
    const int B = hdr[0], Nv = hdr[1], K = hdr[2], ld = hdr[3], iters = hdr[4], hand_face = hdr[5];
    std::vector<float> kp((size_t)B * Nv * K * 3), cams((size_t)Nv * 12), poses((size_t)B * ld), betas((size_t)B * 10);
    if (!rd(f, kp.data(), kp.size() * 4) || !rd(f, cams.data(), cams.size() * 4) || !rd(f, poses.data(), poses.size() * 4) ||
        !rd(f, betas.data(), betas.size() * 4)) { fprintf(stderr, "truncated input\n"); return 1; }
    fclose(f);

    BF(bf_check_device());
    BfModel* model = nullptr;
    BF(bf_model_create(&d, &model));
    if (model->K_used != K) { fprintf(stderr, "K mismatch: model %d, file %d\n", model->K_used, K); return 1; }
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    const int64_t need = bf_workspace_bytes(model, B, Nv, 2, iters);
    if (need <= 0) { fprintf(stderr, "bf_workspace_bytes: %s\n", bf_last_error()); return 3; }
    void* ws = nullptr;
    float *kp_d = nullptr, *poses_d = nullptr, *betas_d = nullptr;
    CK(cudaMalloc(&ws, (size_t)need));
    CK(cudaMalloc(&kp_d, kp.size() * 4)); CK(cudaMalloc(&poses_d, poses.size() * 4)); CK(cudaMalloc(&betas_d, betas.size() * 4));
    CK(cudaMemcpyAsync(kp_d, kp.data(), kp.size() * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(poses_d, poses.data(), poses.size() * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(betas_d, betas.data(), betas.size() * 4, cudaMemcpyHostToDevice, s));
    BfFrames fr;
    BF(bf_frames_bind(model, B, Nv, 2, iters, ws, need, &fr, s));
    BF(bf_pack_keypoints(kp_d, (float*)fr.kp, B, Nv, K, hand_face, nullptr, s));
    CK(cudaMemcpyAsync((void*)fr.cams, cams.data(), cams.size() * 4, cudaMemcpyHostToDevice, s));
    BF(bf_init_theta(model, poses_d, ld, betas_d, fr.theta, B, nullptr, s));
    BF(bf_fit_run(model, &fr, iters, s));
    std::vector<float> theta((size_t)B * model->NP);
    CK(cudaMemcpyAsync(theta.data(), fr.theta, theta.size() * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    FILE* o = fopen(argv[2], "wb");
    if (!o || fwrite(theta.data(), 4, theta.size(), o) != theta.size()) { perror(argv[2]); return 1; }
    fclose(o);
    printf("fit_host: %d frames x %d views, %d iterations, NP %d: ok\n", B, Nv, iters, model->NP);
    BF(bf_model_destroy(model));
    cudaFree(ws); cudaFree(kp_d); cudaFree(poses_d); cudaFree(betas_d);
    return 0;
}
