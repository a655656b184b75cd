// TEST INFRASTRUCTURE ONLY (oracle build). Hand-written equivalent of the file CMake would generate
// from include/gwat/GWATConfig.h.in; the share dir is only read for interpolated PSD files (not used).
#ifndef ORACLE_GWATCONFIG_H
#define ORACLE_GWATCONFIG_H
#define GWAT_VERSION_MAJOR 0
#define GWAT_VERSION_MINOR 0
#define GWAT_ROOT_DIRECTORY "/root/reference/"
#define GWAT_INSTALL_PREFIX "/root/reference/"
#ifndef GWAT_SHARE_DIR
#define GWAT_SHARE_DIR "/root/reference/data/"
#endif
#endif
