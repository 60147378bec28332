// tlb_internal.h -- shared by the host-side translation units of libtoolame_b200.so (not part of the C ABI).
#pragma once
// Record `msg` for tlb_last_error() (per thread) and return `code`.
int tlb_fail(int code, const char *msg);
