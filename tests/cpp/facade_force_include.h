// Test-only.  Force-included (g++ -include) before an UNMODIFIED application source of the reference
// (src/apps/ojph_compress/ojph_compress.cpp, src/apps/ojph_expand/ojph_expand.cpp): OpenJPH's public headers
// come first -- their include guards turn the application's own #includes into no-ops --, then the facade,
// then the class names the application spells are pointed at the facade's classes.  (Aliases rather than
// `#define codestream b200::codestream`, because the applications also call their variable `codestream`.)
#include "ojph_arch.h"
#include "ojph_base.h"
#include "ojph_mem.h"
#include "ojph_file.h"
#include "ojph_params.h"
#include "ojph_codestream.h"
#include "ojph_message.h"
#include "ojph_b200_codestream.hpp"
namespace ojph {
  using codestream_b200 = b200::codestream; using param_siz_b200 = b200::param_siz; using param_cod_b200 = b200::param_cod;
  using param_qcd_b200 = b200::param_qcd; using param_nlt_b200 = b200::param_nlt; using comment_exchange_b200 = b200::comment_exchange;
}
#define codestream codestream_b200
#define param_siz param_siz_b200
#define param_cod param_cod_b200
#define param_qcd param_qcd_b200
#define param_nlt param_nlt_b200
#define comment_exchange comment_exchange_b200
