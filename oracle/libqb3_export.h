/* Stand-in for the export header CMake generates for the reference build
   (QB3lib/CMakeLists.txt:58-62). Used only when compiling oracle/_ref. */
#ifndef LIBQB3_EXPORT_H
#define LIBQB3_EXPORT_H
#define LIBQB3_EXPORT __attribute__((visibility("default")))
#endif
