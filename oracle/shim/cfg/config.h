#define VERSION "5.5.0"
#define PACKAGE_VERSION "5.5.0"
