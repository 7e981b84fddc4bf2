// ciselnik.h -- the reference's numeric constants header (include/lbm_common/ciselnik.h).  Its `no*` / `n*o*` macros cast to `dreal` and
// are used inside the device-side trait classes only, which are tags here; solvers use PI from it.
#pragma once
#ifndef PI
	#define PI 3.1415926535897932384
#endif
