/* srp-b200 host layer -- message callback plumbing and SRPType sizes.
 * Semantics of reference src/utils/message_callback.c:18-35 (format into a 1 KiB
 * buffer, call the user's function if one is installed, otherwise drop) and
 * src/utils/type.c:15-40 (only the unsigned integer and floating types have a size;
 * anything else reports an error and yields 0). */
#include <stdarg.h>
#include <stdio.h>
#include "srp_internal.h"

#define SRP_MESSAGE_CAPACITY 1024

void srpMessage(SRPMessageType type, SRPMessageSeverity severity, const char* sourceFunction,
                const char* format, ...)
{
	if (srpContext.messageCallback.func == NULL)
		return;
	char text[SRP_MESSAGE_CAPACITY];
	va_list ap;
	va_start(ap, format);
	vsnprintf(text, sizeof text, format, ap);
	va_end(ap);
	srpContext.messageCallback.func(type, severity, sourceFunction, text,
	                                srpContext.messageCallback.userParameter);
}

void srpFatalMessage(const char* sourceFunction, const char* format, ...)
{
	char text[SRP_MESSAGE_CAPACITY];
	va_list ap;
	va_start(ap, format);
	vsnprintf(text, sizeof text, format, ap);
	va_end(ap);
	fprintf(stderr, "[srp-b200] %s: %s\n", sourceFunction, text);
	if (srpContext.messageCallback.func != NULL)
		srpContext.messageCallback.func(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, sourceFunction, text,
		                                srpContext.messageCallback.userParameter);
}

size_t srpSizeofType(SRPType type)
{
	switch (type)
	{
		case SRP_UINT8:  return 1;
		case SRP_UINT16: return 2;
		case SRP_UINT32: return 4;
		case SRP_UINT64: return 8;
		case SRP_FLOAT:  return sizeof(float);
		case SRP_DOUBLE: return sizeof(double);
		default:
			srpMessage(SRP_MESSAGE_ERROR, SRP_MESSAGE_SEVERITY_HIGH, __func__, "Unknown type (%i)", type);
			return 0;
	}
}
