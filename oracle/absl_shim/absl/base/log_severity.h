#pragma once
namespace absl { enum class LogSeverity : int { kInfo = 0, kWarning = 1, kError = 2, kFatal = 3 }; }
