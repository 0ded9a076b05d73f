// Minimal stand-in for rapids_logger::level_enum (only the enumerators KMeansParams::verbosity uses).
#pragma once
namespace rapids_logger {
enum class level_enum : int { trace = 0, debug = 1, info = 2, warn = 3, error = 4, critical = 5, off = 6 };
}
