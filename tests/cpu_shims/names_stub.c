/* gj_exif.c names colour spaces in a warning; host_shim.so does not carry gj_common.c (which has the name tables and the
 * CUDA wrappers' users), so the one function is stood in for here.  Test infrastructure. */
#include "../../gpujpeg_b200/csrc/gj_internal.h"
const char* gpujpeg_color_space_get_name(enum gpujpeg_color_space cs) { (void)cs; return "(colour space)"; }
