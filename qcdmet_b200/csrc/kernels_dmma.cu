// kernels_dmma.cu -- placeholder until the site-factorised DMMA kernel lands.
#include "common.cuh"
namespace qcdmet {
struct DmmaPlanData {};
bool dmma_eligible(int, int, int, const int *, const int *, const int *, const int *, const int *,
                   bool, std::string *why) { if (why) *why = "not built"; return false; }
int dmma_plan_build(DmmaPlanData **out, int, int, int, int, const int *, const int *, const int *,
                    const int *, const int *) { *out = nullptr; return QCDMET_OK; }
void dmma_plan_free(DmmaPlanData *) {}
int dmma_prepare_orbitals(DmmaPlanData *, const JacView &, cudaStream_t) { return QCDMET_OK; }
int launch_jacobian_dmma(DmmaPlanData *, const JacView &, cudaStream_t, int *, double *, cudaEvent_t, cudaEvent_t)
{ set_error("DMMA kernel not built"); return QCDMET_E_STATE; }
}
