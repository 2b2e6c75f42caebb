// TEMPORARY: render backward lands in vb_render_bwd.cu.
#include "vb_common.cuh"
extern "C" size_t vb200_render_bwd_workspace(const VbGrid*, int) { return 0; }
extern "C" int vb200_render_bwd(const VbGrid*, const VbTables*, const float*, const VbRenderIn*, int,
                                const VbRenderOut*, const VbRenderGrad*, int, void*, size_t, void*) {
  return VB200_ERR_ARG;
}
