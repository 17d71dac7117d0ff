/* ncurses is absent; status.c:19 includes it but the functions the oracle links use none of it. */
#ifndef KA9Q_ORACLE_NCURSES_SHIM_H
#define KA9Q_ORACLE_NCURSES_SHIM_H 1
#endif
