#define HAVE_PTHREAD_BARRIER_WAIT 1
#define HAVE_PTHREAD_ATTR_SETAFFINITY_NP 1
#define HAVE_LINUX_PERF_EVENT_H 1
#define PACKAGE_STRING "multicore-hashjoins (oracle build)"
