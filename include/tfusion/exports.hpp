#pragma once
#if defined(__GNUC__)
#define KF_EXPORTS __attribute__((visibility("default")))
#else
#define KF_EXPORTS
#endif
