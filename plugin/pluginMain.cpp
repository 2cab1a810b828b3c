/*
 * Plugin entry point of libosm_b200_plugin.so: the symbol the reference's plugin loader looks up with dlsym
 * (src/core/componentManager.cpp:253-264) and calls once per registration round with (confman, compman, iteration)
 * (:381-418; the registerFunction type, src/include/core/componentManager.hpp:23).  It returns the linked list of
 * sComponentInfo records of the components this library adds (the SDK's sample: plugindev/pluginMain.cpp:43-58).
 */
#include <core/smileCommon.hpp>
#include <core/componentManager.hpp>

#include "lldBlockB200.hpp"

static const registerFunction b200_components[] = {
  cLldBlockB200::registerComponent,
  NULL
};

extern "C" __attribute__((visibility("default")))
sComponentInfo *registerPluginComponent(cConfigManager *confman, cComponentManager *compman, int iteration)
{
  sComponentInfo *head = NULL, *tail = NULL;
  for (int i = 0; b200_components[i] != NULL; i++) {
    sComponentInfo *cur = (b200_components[i])(confman, compman, iteration);
    if (cur == NULL) continue;
    cur->next = NULL;
    if (head == NULL) head = cur; else tail->next = cur;
    tail = cur;
  }
  return head;
}
