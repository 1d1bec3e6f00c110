/* Forwarding header: the declarations the reference keeps in include/srp/message_callback.h live in srp/api.h. */
#pragma once
#include "srp/api.h"
