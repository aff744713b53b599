/* empty stand-in for <conio.h> (test infrastructure) */
