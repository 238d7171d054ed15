// api_macros.h -- error plumbing shared by the translation units that define C-ABI entry points.
#pragma once
#include "gsb_internal.h"

namespace gsb {
int set_error(gsb_ctx_t ctx, const std::string &msg);
const std::string &last_error();
bool ctx_alive(gsb_ctx_t ctx);
}  // namespace gsb

#define GSB_NULLCHK(h)                                         \
  if (!(h)) {                                                 \
    gsb::set_error(nullptr, "NULL handle passed to libgsb200"); \
    return GSB_EINVAL;                                        \
  }
#define API_BEGIN try {
#define API_END(ctx)                               \
  }                                                \
  catch (const gsb::Error &e) {                    \
    gsb::set_error((ctx), e.msg);                  \
    return e.code;                                 \
  }                                                \
  catch (const std::exception &e) {                \
    gsb::set_error((ctx), e.what());               \
    return GSB_EINVAL;                             \
  }                                                \
  return GSB_OK;


