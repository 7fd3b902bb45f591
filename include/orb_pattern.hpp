// orb_pattern.hpp -- same declaration as the reference's include/orb_pattern.hpp:2; defined in liborb_b200.so.
//x1,y1,x2,y2
#ifdef __cplusplus
extern "C" {
#endif
extern int bit_pattern_31_[256*4];
#ifdef __cplusplus
}
#endif
