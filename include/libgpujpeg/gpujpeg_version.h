/* Forwarder: the whole libgpujpeg C ABI of this build lives in ../gpujpeg_b200.h
 * (replaces the reference's libgpujpeg/gpujpeg_version.h). */
#ifndef GPUJPEG_VERSION_H_FWD_B200
#define GPUJPEG_VERSION_H_FWD_B200
#include "../gpujpeg_b200.h"
#endif
