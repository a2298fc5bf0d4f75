#define BUILD_B200 1
