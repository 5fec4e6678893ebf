/* connect_local.c -- LD_PRELOAD interposer for running the reference's TCP sender on one machine (test infrastructure).
 *
 * multiple_connections_network_client_sender.c:47 connects to a literal server address (10.1.212.25) on PORT + i.  The
 * sender is compiled unmodified from where it lies (oracle/Makefile: ref); this shim sends its connect() calls to
 * 127.0.0.1 instead and, when FR_SENDER_PORT_FROM / FR_SENDER_PORT_TO are set, shifts the port range (8080.. may be taken).
 */
#define _GNU_SOURCE
#include <arpa/inet.h>
#include <dlfcn.h>
#include <netinet/in.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>

int connect(int fd, const struct sockaddr* addr, socklen_t len) {
  static int (*real_connect)(int, const struct sockaddr*, socklen_t);
  if (!real_connect) real_connect = (int (*)(int, const struct sockaddr*, socklen_t))dlsym(RTLD_NEXT, "connect");
  if (addr && addr->sa_family == AF_INET && len >= (socklen_t)sizeof(struct sockaddr_in)) {
    struct sockaddr_in a;
    memcpy(&a, addr, sizeof a);
    a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
    const char *from = getenv("FR_SENDER_PORT_FROM"), *to = getenv("FR_SENDER_PORT_TO");
    if (from && to) a.sin_port = htons((unsigned short)(ntohs(a.sin_port) - atoi(from) + atoi(to)));
    return real_connect(fd, (const struct sockaddr*)&a, sizeof a);
  }
  return real_connect(fd, addr, len);
}
