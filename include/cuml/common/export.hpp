// Symbol visibility for the ML:: C++ surface (role of reference cpp/include/cuml/common/export.hpp:8-14).
#pragma once
#define CUML_EXPORT __attribute__((visibility("default")))
