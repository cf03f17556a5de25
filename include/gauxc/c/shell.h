/* Shim: the declarations of the reference's <gauxc/c/shell.h> live in gauxc_b200.h, so a client written
 * against the reference's C API compiles with -I<this repo>/include unchanged. */
#pragma once
#include "../../gauxc_b200.h"
