/* Minimal OpenGL type stub: this image has no GL headers, and the reference's cudamanager includes <cudaGL.h> (which includes <GL/gl.h>) for
 * its graphics-interop entry points.  Only the types cudaGL.h mentions are needed to COMPILE the reference's GPU plugin (oracle/ref_gpu_build.mk);
 * no GL function is ever called by the benchmark. */
#ifndef PXB_GL_STUB_H
#define PXB_GL_STUB_H
typedef unsigned int GLenum;
typedef unsigned int GLuint;
typedef int GLint;
typedef int GLsizei;
typedef unsigned char GLboolean;
typedef float GLfloat;
typedef void GLvoid;
#endif
