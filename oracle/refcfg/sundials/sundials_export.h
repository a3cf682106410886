/* TEST INFRASTRUCTURE (oracle/_ref build) -- stands in for the cmake-generated export header. */
#ifndef SUNDIALS_EXPORT_H
#define SUNDIALS_EXPORT_H
#define SUNDIALS_EXPORT __attribute__((visibility("default")))
#define SUNDIALS_NO_EXPORT __attribute__((visibility("hidden")))
#define SUNDIALS_DEPRECATED __attribute__((__deprecated__))
#define SUNDIALS_DEPRECATED_EXPORT SUNDIALS_EXPORT SUNDIALS_DEPRECATED
#define SUNDIALS_DEPRECATED_NO_EXPORT SUNDIALS_NO_EXPORT SUNDIALS_DEPRECATED
#endif
