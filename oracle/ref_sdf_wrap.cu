// TEST INFRASTRUCTURE -- C entry point around the reference's own host launcher sdf_cuda()
// (sdf/sdf/csrc/sdf_cuda_kernel.cu:307-335, compiled unchanged next to this file by oracle/build_ref_sdf.sh).
// The tensors are described the way sdf/sdf/sdf.py:12-15 and code/utils/fitting.py:367-368 hand them over:
//   phi [B,G,G,G] (written in place), faces int32 with faces.size(0) = faces_dim0, vertices [B,num_verts,3].
// The reference call passes faces as [1,F,3], i.e. faces_dim0 = 1; faces_dim0 = F is the intended use.
#include <ATen/ATen.h>
#include <cuda_runtime.h>

at::Tensor sdf_cuda(at::Tensor phi, at::Tensor faces, at::Tensor vertices);      // defined in the reference file

extern "C" int ref_sdf_cuda(void* phi_dev, const int* faces_dev, const void* verts_dev, int batch, int grid, int faces_dim0,
                            int num_verts, int is_double) {
    at::Tensor phi, faces, verts;
    phi.ptr = phi_dev; phi.sizes[0] = batch; phi.sizes[1] = grid; phi.sizes[2] = grid; phi.sizes[3] = grid;
    phi.st = is_double ? at::ScalarType::Double : at::ScalarType::Float;
    faces.ptr = const_cast<int*>(faces_dev); faces.sizes[0] = faces_dim0; faces.sizes[1] = 3; faces.st = at::ScalarType::Int;
    verts.ptr = const_cast<void*>(verts_dev); verts.sizes[0] = batch; verts.sizes[1] = num_verts; verts.sizes[2] = 3;
    verts.st = phi.st;
    sdf_cuda(phi, faces, verts);                           // launches on the legacy default stream (:321)
    return (int)cudaDeviceSynchronize();
}
