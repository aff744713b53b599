/* minimal stand-in for <windows.h> so the unmodified version1 sources compile under gcc (test infrastructure) */
#ifndef B2_WINDOWS_SHIM_H
#define B2_WINDOWS_SHIM_H
#include <sys/timeb.h>
#include <stdlib.h>
#include <string.h>
typedef unsigned long DWORD;
typedef int BOOL;
typedef void *HANDLE;
typedef unsigned short WORD;
typedef unsigned char BYTE;
#ifndef min
#define min(a,b) (((a)<(b))?(a):(b))
#endif
#ifndef max
#define max(a,b) (((a)>(b))?(a):(b))
#endif
#define _ftime ftime
#define _timeb timeb
#endif
