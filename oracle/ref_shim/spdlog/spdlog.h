// TEST INFRASTRUCTURE ONLY: stand-in so that reference sources which merely include spdlog parse here.
// Logging is provided by oracle/ref_shim/common/Error.hh (a null logger).
#pragma once
