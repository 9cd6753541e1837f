// Test-only stand-in for the g2o type headers: include/Converter.h only names these types in declarations.
#pragma once
namespace g2o {
#ifndef COEB_REF_SHIM_G2O_TYPES
#define COEB_REF_SHIM_G2O_TYPES
struct SE3Quat {};
struct Sim3 {};
#endif
}  // namespace g2o
