/* Empty stand-in for <sys/sysctl.h>, which glibc >= 2.32 no longer ships.
 * The reference includes it unconditionally (src/core/parallel.cpp:38) but on
 * Linux only uses sysconf() (src/core/parallel.cpp:876-877). Test infrastructure
 * for building oracle/_ref; not part of the product. */
