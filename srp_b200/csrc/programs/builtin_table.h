/* srp-b200 built-in programs -- the program table: X(deviceId, name, vertex shader,
 * fragment shader, uniform type).  Consumed by builtin_host.c (name -> host function
 * pointers) and builtin_device.cu (device dispatch + registration). */
#define SRPB_PROGRAM_TABLE(X) \
	X(0, texcube,  srpb_texcube_vs, srpb_texcube_fs,  SrpbTexCubeUniform) \
	X(1, gouraud,  srpb_gouraud_vs, srpb_gouraud_fs,  SrpbGouraudUniform) \
	X(2, vcolor,   srpb_vcolor_vs,  srpb_gouraud_fs,  SrpbTransform) \
	X(3, primid,   srpb_primid_vs,  srpb_primid_fs,   SrpbTransform) \
	X(4, tagged,   srpb_tagged_vs,  srpb_tagged_fs,   SrpbTransform) \
	X(5, solid,    srpb_solid_vs,   srpb_solid_fs,    SrpbSolidUniform) \
	X(6, depthout, srpb_vcolor_vs,  srpb_depthout_fs, SrpbTransform) \
	X(7, mixed,    srpb_mixed_vs,   srpb_mixed_fs,    SrpbTransform)
