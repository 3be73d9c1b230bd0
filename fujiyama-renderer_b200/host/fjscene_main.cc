// fjscene — command-line front end: `fjscene file.scn` executes a scene-description file on the GPU,
// like the reference's bin/scene (tools/scene_parser/main.cc:9-52).  `-` or no argument reads stdin.
#include "fjscene.h"

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <string>

int main(int argc, char **argv) {
  if (argc == 2 && strcmp(argv[1], "--help") == 0) {
    printf("Usage: fjscene [file.scn]\n  FJ_DEVICE=<ordinal> selects the GPU, FJ_SEED=<n> the stochastic-shader seed.\n");
    return 0;
  }
  const char *dev = getenv("FJ_DEVICE");
  if (dev) fjscene_set_device(atoi(dev), 0, 1);
  fjscene_parser *p = fjscene_parser_new();
  int rc = 0;
  if (argc >= 2 && strcmp(argv[1], "-") != 0) rc = fjscene_parse_file(p, argv[1]);
  else {
    std::string line;
    while (std::getline(std::cin, line)) if ((rc = fjscene_parse_line(p, line.c_str()))) break;
  }
  fjscene_parser_free(p);
  return rc ? 1 : 0;
}
