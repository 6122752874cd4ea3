// placeholder, replaced below
#include "hd_internal.h"
namespace hd
{
  bool fast6d_supported(const hd_advection *) { return false; }
  int  launch_fast6d(hd_advection *, void *, const void *, const void *, double, const FusedUpdate &) { return fail(HD_ERR_UNSUPPORTED, "fast6d not built"); }
  void fast6d_release(hd_advection *) {}
}
