// forwards to the stand-in (oracle/ref_stubs/b2a_ros_stub.h); test infrastructure only
#include "b2a_ros_stub.h"
