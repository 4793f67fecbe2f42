/* oracle/ref_o4_shim.cu -- TEST INFRASTRUCTURE ONLY (oracle "O4").  metrans' NvCodec sources log through a global
 * `simplelogger::Logger *logger` that every metrans application defines (e.g. app/AppMeTrans/AppMeTrans.cpp);
 * the oracle library defines it the same way. */
#include "NvCommon.h"
simplelogger::Logger *logger = simplelogger::LoggerFactory::CreateConsoleLogger();
