/* oracle/ref_compat: the reference includes highgui but uses nothing from it on this path. */
#ifndef SDORB_REF_COMPAT_HIGHGUI_HPP
#define SDORB_REF_COMPAT_HIGHGUI_HPP
#include "../core/core.hpp"
#endif
