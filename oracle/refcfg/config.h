/* Build-time configuration for compiling the reference x264 snapshot in place
 * (oracle/_ref).  The reference's ./configure would emit the same four lines
 * (S/configure:411-417, S/version.sh:13-19 for a tree without git history).
 * TEST INFRASTRUCTURE ONLY. */
#define X264_VERSION ""
#define X264_POINTVER "0.66.x"
#define fseek fseeko
#define ftell ftello
