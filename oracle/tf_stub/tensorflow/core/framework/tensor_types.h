// see oracle/tf_stub/tf_stub.h (stand-in for the TensorFlow op framework; test infrastructure)
#include "tf_stub.h"
