/*
 * label_model.c -- TEST INFRASTRUCTURE: CPU model of k_label (odr_audioenc_b200/csrc/mp2_kernels.cu) checked against
 * the oracle's verbatim psy-1 labelling (psy1_tonal / psy1_noise / psy1_subsample in oracle/mp2_oracle.c, which this file
 * includes) on real signal frames: mask-driven tonal walk, noise maskers and decimation on a shared next[] array,
 * including the frames where the tonal and the noise list run into each other.  Prints "bad N".
 * usage: label_model PCM_FILE(s16 stereo 48 kHz) N_FRAMES
 */
#include "../oracle/mp2_oracle.c"
#include <stdio.h>
#define MAX_TONAL 104
static int tonal_run(int i){ if(i<3||i>500) return 0; if(i<63) return 2; if(i<127) return 3; if(i<255) return 6; return 12;}
static int next_bit(const unsigned*m,int p){ int w=(p+1)>>5; if(w>=16) return L_LAST; unsigned bits=m[w]&(~0u<<((p+1)&31)); while(!bits){ if(++w>=16) return L_LAST; bits=m[w];} return w*32+__builtin_ffs(bits)-1;}
typedef struct { double t_x[MAX_TONAL], n_x[28]; int t_part[MAX_TONAL], n_part[28]; int n_tone, n_noise; } maskers;
static void gpu_style(double *x, const double *wgt, int fq, const int *map, maskers *out)
{
    const double *hear = MP2_LTG_HEAR[fq], *bark = MP2_LTG_BARK[fq];
    unsigned cand[16]={0}, t0[16]={0}, tone_mask[16]={0};
    for (int i=0;i<512;i++){ int peak = i>=2 && i<500 && x[i]>x[i-1] && x[i]>=x[i+1]; int pass=peak; if(peak){ int run=tonal_run(i); double mx=x[i]-7; for(int j=2;j<=run;j++) if(mx<x[i-j]||mx<x[i+j]){pass=0;break;} } if(peak) cand[i>>5]|=1u<<(i&31); if(pass) t0[i>>5]|=1u<<(i&31);} 
#define X(j) x[(j)]
#define TONE_BIT(j) ((tone_mask[(j)>>5]>>((j)&31))&1)
    short next[512];
    for (int i=0;i<512;i++) next[i]=L_STOP;
    int tone=L_LAST,last=L_LAST,last_but_one=L_LAST,mod_end=-1;
    for (int c = next_bit(cand,-1); c!=L_LAST;) {
        const int run=tonal_run(c); int tonal;
        if (c-run>mod_end) tonal=(t0[c>>5]>>(c&31))&1;
        else { tonal=1; const double mx=X(c)-7; for(int j=2;j<=run;j++) if(mx<X(c-j)||mx<X(c+j)){tonal=0;break;} }
        if(!tonal){ c=next_bit(cand,c); continue; }
        if(tone==L_LAST) tone=c;
        if(last!=L_LAST) next[last]=(short)c;
        const int beyond=next_bit(cand,c+run);
        next[c]=(short)beyond;
        if((c-last)<=run){ if(last_but_one!=L_LAST) next[last_but_one]=(short)c; }
        if(c>1&&c<500){ const double tmp=add_db(X(c-1),X(c+1)); X(c)=add_db(X(c),tmp); }
        for(int j=1;j<=run;j++){ X(c-j)=DBMIN; X(c+j)=DBMIN; next[c-j]=next[c+j]=L_STOP; tone_mask[(c-j)>>5]&=~(1u<<((c-j)&31)); }
        tone_mask[c>>5]|=1u<<(c&31);
        mod_end=c+run; last_but_one=last; last=c; c=beyond;
    }
    if(last!=L_LAST) next[last]=L_LAST;
    const int *cbound=MP2_CBOUND[fq]; const int ncb=MP2_CB_COUNT[fq]-1;
    int noise=L_LAST,last_n=L_LAST;
    for (int b=0;b<ncb;b++){ int c0=cbound[b],c1=cbound[b+1]; double weight=0.0,sum=DBMIN;
        for(int j=c0;j<c1;j++){ if(!TONE_BIT(j) && x[j]!=DBMIN){ sum=add_db(x[j],sum); weight+=wgt[j]; x[j]=DBMIN; } }
        int centre; if(sum<=DBMIN) centre=(c1+c0)/2; else { double index=weight*pow(10.0,-0.1*sum); centre=c0+(int)(index*(double)(c1-c0)); }
        if(TONE_BIT(centre)){ if(TONE_BIT(centre+1)) centre++; else centre--; }
        if(last_n==L_LAST) noise=centre; else { next[centre]=L_LAST; next[last_n]=(short)centre; }
        X(centre)=sum; tone_mask[centre>>5]&=~(1u<<(centre&31)); last_n=centre; }
    for(int pass=0;pass<2;pass++){ int head= pass==0?tone:noise; int i=head,old=L_STOP;
        for(int g=0;i!=L_LAST&&i!=L_STOP&&g<600;g++){ if(X(i)<hear[map[i]]){ X(i)=DBMIN; if(old==L_STOP) head=next[i]; else next[old]=next[i]; } else old=i; i=next[i]; }
        if(pass==0) tone=head; else noise=head; }
    { int i=tone,old=L_STOP; for(int g=0;i!=L_LAST&&i!=L_STOP&&g<600;g++){ const int nx=next[i]; if(nx==L_LAST||nx==L_STOP) break;
        if(bark[map[nx]]-bark[map[i]]<0.5){ if(X(nx)>X(i)){ if(old==L_STOP) tone=nx; else next[old]=(short)nx; X(i)=DBMIN; i=nx; } else { X(nx)=DBMIN; next[i]=next[nx]; old=i; } } else { old=i; i=nx; } } }
    int n_tone=0; for(int k=tone;k!=L_LAST&&k!=L_STOP&&n_tone<MAX_TONAL;k=next[k]){ out->t_x[n_tone]=X(k); out->t_part[n_tone]=map[k]; n_tone++; }
    int n_noise=0; for(int k=noise;k!=L_LAST&&k!=L_STOP&&n_noise<28;k=next[k]){ out->n_x[n_noise]=X(k); out->n_part[n_noise]=map[k]; n_noise++; }
    out->n_tone=n_tone; out->n_noise=n_noise;
}
int main(int argc,char**argv){
    const char*path=argv[1]; long nfr=atol(argv[2]);
    FILE*f=fopen(path,"rb"); int16_t*pcm=malloc(nfr*1152*2*2); fread(pcm,2,nfr*1152*2,f); fclose(f);
    mp2o_cfg c; mp2o_configure(&c,48000,'j',192,1,0);
    int fq=c.psy_freq; static psy1_lines P; int bad=0;
    for(long fr=0;fr<nfr;fr++) for(int ch=0;ch<2;ch++){
        double xr[1024],energy[513],wgt[512]={0},xg[512];
        psy1_make_map(fq,P.map);
        for(int i=0;i<1024;i++) xr[i]=pcm_at(pcm,2,ch,fr*1152-192+i)*MP2_HANN[i];
        mp2o_fht1024(xr);
        energy[0]=xr[0]*xr[0]; for(int i=1;i<512;i++) energy[i]=(xr[i]*xr[i]+xr[1024-i]*xr[1024-i])/2.0; energy[512]=xr[512]*xr[512];
        for(int i=0;i<512;i++){ P.x[i]= energy[i]<1E-20? -200.0+POWERNORM : 10*log10(energy[i])+POWERNORM; P.next[i]=L_STOP; P.type[i]=0; xg[i]=P.x[i]; }
        const int*cb=MP2_CBOUND[fq]; for(int b=0;b<MP2_CB_COUNT[fq]-1;b++) for(int j=cb[b];j<cb[b+1];j++) wgt[j]= 1073741824*energy[j]*(double)(j-cb[b])/(double)(cb[b+1]-cb[b]);
        int tone,noise; psy1_tonal(&P,&tone); psy1_noise(&P,&noise,fq,energy); psy1_subsample(&P,fq,&tone,&noise);
        maskers M; gpu_style(xg,wgt,fq,P.map,&M);
        /* compare lists */
        int n=0,ok=1; for(int t=tone;t!=L_LAST&&t!=L_STOP;t=P.next[t],n++){ if(n>=M.n_tone||M.t_x[n]!=P.x[t]||M.t_part[n]!=P.map[t]) ok=0; } if(n!=M.n_tone) ok=0;
        int m=0; for(int t=noise;t!=L_LAST&&t!=L_STOP;t=P.next[t],m++){ if(m>=M.n_noise||M.n_x[m]!=P.x[t]||M.n_part[m]!=P.map[t]) ok=0; } if(m!=M.n_noise) ok=0;
        if(!ok){ if(bad<6){ printf("frame %ld ch %d: oracle tone %d noise %d | gpu-style tone %d noise %d\n",fr,ch,n,m,M.n_tone,M.n_noise);
            printf("  oracle tonal:"); for(int t=tone;t!=L_LAST&&t!=L_STOP;t=P.next[t]) printf(" %d(%.2f)",t,P.x[t]); printf("\n  oracle noise:"); for(int t=noise;t!=L_LAST&&t!=L_STOP;t=P.next[t]) printf(" %d(%.2f)",t,P.x[t]);
            printf("\n  gpu tonal parts:"); for(int k=0;k<M.n_tone;k++) printf(" p%d(%.2f)",M.t_part[k],M.t_x[k]); printf("\n  gpu noise parts:"); for(int k=0;k<M.n_noise;k++) printf(" p%d(%.2f)",M.n_part[k],M.n_x[k]); printf("\n"); }
            bad++; }
    }
    printf("bad %d\n",bad); return 0; }
